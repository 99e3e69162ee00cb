"""`torch_geometric.loader.DataLoader` (oracle shim, test infrastructure): a torch DataLoader whose
collate function is `Batch.from_data_list`; shuffle / drop_last=False semantics are torch's
(reference dss2_run.py:68-69)."""
import torch.utils.data

from ..data import Batch


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)
