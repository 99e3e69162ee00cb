"""The dense head shared by GAT_DSSE / GINE_DSSE / gnn_dsse (reference networks.py:62-63,104-105,150-151): Linear(dim_feat, dim_dense),
Linear(dim_dense, dim_out) with nothing in between.  For the script's sizes (8 -> 32 -> 2) the kernels never put the hidden row or its
gradient in memory (csrc/gat.cu k_mlp2_bwd_nh); other sizes keep the stored-h kernels."""
from . import _lib


def head_fwd(lib, n, x, din, w1, b1, dmid, w2, b2, dout, h, out, st):
    nh = bool(lib.dss2_mlp2_nh_supported(din, dmid, dout))
    _lib.check(lib.dss2_mlp2_fwd(n, x, din, w1, b1, dmid, w2, b2, dout, None if nh else h, out, st), "dss2_mlp2_fwd")


def head_bwd(lib, n, x, din, w1, b1, dmid, w2, dout, h, grad_out, gh, gx, partials, pstride, st):
    """partials: pointer at the head's first column ([w1 | b1 | w2 | b2]) of the per-CTA partial rows."""
    if lib.dss2_mlp2_nh_supported(din, dmid, dout):
        _lib.check(lib.dss2_mlp2_bwd_nh(n, x, din, w1, b1, dmid, w2, dout, grad_out, gx, partials, pstride, st), "dss2_mlp2_bwd_nh")
    else:
        _lib.check(lib.dss2_mlp2_bwd(n, x, din, w1, dmid, w2, dout, h, grad_out, gh, gx, partials, pstride, st), "dss2_mlp2_bwd")
