"""`torch_geometric.nn` subset (oracle shim, test infrastructure): the conv classes and `Sequential`
(reference networks.py:4,55,111,153 build the non-PowerFlowNet models with it)."""
from torch import nn as _nn

from . import conv
from .conv import (GCN2Conv, FAConv, TAGConv, GINEConv, MessagePassing, GCNConv, ChebConv,
                   GATv2Conv, gcn_norm)


class Sequential(_nn.Module):
    """PyG `nn.Sequential(input_args, modules)`: `input_args` names the forward arguments ('x, edge_index, edge_attr'); an entry
    `(module, 'a, b -> c')` calls module(a, b) and binds the result to c, a bare module is applied to the previous result.
    Children are registered as `module_{i}` (that is what names the parameters in a state_dict)."""

    def __init__(self, input_args, modules):
        super().__init__()
        self._inputs = [a.strip() for a in input_args.split(",")]
        self._specs = []
        for i, entry in enumerate(modules):
            if isinstance(entry, (tuple, list)):
                mod, desc = entry
                ins, outs = desc.split("->")
                ins = [a.strip() for a in ins.split(",")]
                outs = [a.strip() for a in outs.split(",")]
            else:
                mod, ins, outs = entry, None, None
            self.add_module(f"module_{i}", mod)
            self._specs.append((f"module_{i}", ins, outs))

    def forward(self, *args):
        env = dict(zip(self._inputs, args))
        last = args[0]
        prev_outs = [self._inputs[0]]
        for name, ins, outs in self._specs:
            mod = getattr(self, name)
            if ins is None:
                # PyG: a module given WITHOUT a signature string operates on the output of its preceding module AND REBINDS it
                # (in_desc = out_desc = the previous call's out_desc), so `(conv, 'x, ei -> x'), LeakyReLU()` feeds the activated x
                # to the next conv.  (Round 1 of this shim dropped the rebinding: the activations between GATv2 / GINE layers had
                # no effect on the following layer - found by running the reference's script against the drop-in modules.)
                last = mod(last)
                if len(prev_outs) == 1:
                    env[prev_outs[0]] = last
            else:
                res = mod(*[env[k] for k in ins])
                if len(outs) == 1:
                    env[outs[0]] = res
                else:
                    for k, v in zip(outs, res):
                        env[k] = v
                last = res
                prev_outs = outs
        return last
