"""dss2_b200 internals: ctypes binding (`_lib`), batch structure (`graph`), launch sequences and autograd
bridges (`ops`), scenario store / dataset builder (`dataset`), synthetic grids (`synth`), CUDA batch packer
(`batching`) and the CUDA-graph trainer (`trainer`).  The drop-in surface for the reference's script is the
three top-level modules next to this package: `networks`, `data`, `loadsampling`."""
