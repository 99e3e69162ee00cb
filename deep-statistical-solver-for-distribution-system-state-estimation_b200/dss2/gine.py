"""Host side of GINE_DSSE (reference networks.py:71-111): spec, flat parameter layout under the names
named_parameters() reports (`nn.*` once: the ONE Linear shared by all GINEConv layers; `model.module_{2l}.lin.*`; the head), launch
sequence, autograd bridge.  All arithmetic is in csrc/gat.cu; there is no CPU fallback."""
import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from .head import head_bwd, head_fwd
from .ops import ParamPack, _align4, require_cuda, resolve_graph, stage_rows

GINE_C = 8


@dataclass(frozen=True)
class GINESpec:
    dim_feat: int
    dim_dense: int
    dim_out: int
    num_layers: int
    edge_dim: int
    eps: float = 0.0
    act_slope: float = 0.01
    train_eps: bool = False   # eps of every GINEConv is a parameter `model.module_{2l}.eps` [1] (PyG GINEConv(train_eps=True))

    @property
    def n_conv(self):
        return self.num_layers - 1

    def layout(self):
        """name -> (offset, numel) of the PARAMETERS; gradient buffers are `grad_size()` long: behind the parameters sits one
        [nn.weight | nn.bias] slot per layer for that layer's share of the shared Linear's gradient."""
        off, table = 0, {}

        def put(name, n):
            nonlocal off
            table[name] = (off, n)
            off += n

        c = self.dim_feat
        put("nn.weight", c * c)
        put("nn.bias", c)
        for l in range(self.n_conv):
            put(f"model.module_{2 * l}.lin.weight", c * self.edge_dim)
            put(f"model.module_{2 * l}.lin.bias", c)
            if self.train_eps:   # directly behind lin.bias: the backward kernel writes its gradient as the next element of that partial row
                put(f"model.module_{2 * l}.eps", 1)
            off = _align4(off)
        i = 2 * self.n_conv
        put(f"model.module_{i}.weight", self.dim_dense * c)
        put(f"model.module_{i}.bias", self.dim_dense)
        put(f"model.module_{i + 1}.weight", self.dim_out * self.dim_dense)
        put(f"model.module_{i + 1}.bias", self.dim_out)
        return table, _align4(off)

    def nn_share_offset(self, l):
        return self.layout()[1] + l * (self.dim_feat * self.dim_feat + self.dim_feat)

    def grad_size(self):
        return _align4(self.nn_share_offset(self.n_conv))


def validate_gine_spec(sp):
    if sp.dim_feat != GINE_C:
        raise NotImplementedError(f"GINE_DSSE kernels are built for dim_feat == {GINE_C}, got {sp.dim_feat}")
    if not (1 <= sp.edge_dim <= 8 and 1 <= sp.dim_dense <= 32 and 1 <= sp.dim_out <= 8 and sp.num_layers >= 2):
        raise NotImplementedError(f"GINE_DSSE kernels support edge_dim <= 8, dim_dense <= 32, dim_out <= 8, num_layers >= 2; got {sp}")


class GINERunner:
    def __init__(self, spec):
        validate_gine_spec(spec)
        self.spec = spec
        self.table, self.flat_size = spec.layout()
        self.grad_size = spec.grad_size()
        self.lib = _lib.load()
        self.num_partials = self.lib.dss2_num_partials()

    def _p(self, flat, name):
        return ctypes.c_void_p(flat.data_ptr() + 4 * self.table[name][0])

    def _layer(self, flat, l):
        return [self._p(flat, "nn.weight"), self._p(flat, "nn.bias"), self._p(flat, f"model.module_{2 * l}.lin.weight"),
                self._p(flat, f"model.module_{2 * l}.lin.bias")]

    def _eps(self, flat, l):
        return self._p(flat, f"model.module_{2 * l}.eps") if self.spec.train_eps else None

    def alloc(self, num_nodes, device, need_grad=True):
        sp = self.spec
        f32 = dict(dtype=torch.float32, device=device)
        b = {"acts": torch.empty(sp.n_conv, num_nodes, GINE_C, **f32), "h": torch.empty(num_nodes, sp.dim_dense, **f32),
             "out": torch.empty(num_nodes, sp.dim_out, **f32)}
        if need_grad:
            b["g8"] = [torch.empty(num_nodes, GINE_C, **f32) for _ in range(2)]
            b["gh"] = torch.empty(num_nodes, sp.dim_dense, **f32)
            b["ws"] = torch.empty(self.lib.dss2_gine_ws_bytes(num_nodes) // 4, **f32)
            b["partials"] = torch.zeros(self.num_partials, self.grad_size, **f32)
        return b

    def forward(self, graph, x, xs, ea, eas, flat, bufs):
        sp, lib, st = self.spec, self.lib, _lib.stream()
        for l in range(sp.n_conv):
            xin, stride = (x, xs) if l == 0 else (bufs["acts"][l - 1], GINE_C)
            _lib.check(lib.dss2_gine_fwd_ex(graph.ref, _lib.ptr(xin), stride, _lib.ptr(ea), eas, sp.edge_dim, *self._layer(flat, l), sp.eps,
                                            self._eps(flat, l), 1, sp.act_slope, _lib.ptr(bufs["acts"][l]), st), "dss2_gine_fwd")
        i = 2 * sp.n_conv
        head_fwd(lib, graph.num_nodes, _lib.ptr(bufs["acts"][sp.n_conv - 1]), GINE_C, self._p(flat, f"model.module_{i}.weight"),
                 self._p(flat, f"model.module_{i}.bias"), sp.dim_dense, self._p(flat, f"model.module_{i + 1}.weight"),
                 self._p(flat, f"model.module_{i + 1}.bias"), sp.dim_out, _lib.ptr(bufs["h"]), _lib.ptr(bufs["out"]), st)
        return bufs["out"]

    def backward(self, graph, x, xs, ea, eas, flat, bufs, grad_out, flat_grad, need_gx=False):
        """grad_out [Nt, dim_out] -> flat_grad [grad_size]: parameters first, then the per-layer shares of the shared Linear, which a
        final fixed-order column sum folds into the `nn.*` slot."""
        sp, lib, st = self.spec, self.lib, _lib.stream()
        part, pstride = bufs["partials"], self.grad_size

        def pp(offset):
            return ctypes.c_void_p(part.data_ptr() + 4 * offset)

        i = 2 * sp.n_conv
        gy = bufs["g8"][0]
        head_bwd(lib, graph.num_nodes, _lib.ptr(bufs["acts"][sp.n_conv - 1]), GINE_C, self._p(flat, f"model.module_{i}.weight"),
                 self._p(flat, f"model.module_{i}.bias"), sp.dim_dense, self._p(flat, f"model.module_{i + 1}.weight"), sp.dim_out,
                 _lib.ptr(bufs["h"]), _lib.ptr(grad_out), _lib.ptr(bufs["gh"]), _lib.ptr(gy), pp(self.table[f"model.module_{i}.weight"][0]), pstride, st)
        gx_out = None
        for l in reversed(range(sp.n_conv)):
            xin, stride = (x, xs) if l == 0 else (bufs["acts"][l - 1], GINE_C)
            want_gx = l > 0 or need_gx
            gx = bufs["g8"][(sp.n_conv - l) & 1] if want_gx else None
            _lib.check(lib.dss2_gine_bwd_ex(graph.ref, _lib.ptr(xin), stride, _lib.ptr(ea), eas, sp.edge_dim, *self._layer(flat, l), sp.eps,
                                            self._eps(flat, l), 1, sp.act_slope, _lib.ptr(bufs["acts"][l]), _lib.ptr(gy), _lib.ptr(gx), _lib.ptr(bufs["ws"]),
                                         bufs["ws"].numel() * 4, pp(self.table[f"model.module_{2 * l}.lin.weight"][0]),
                                         pp(sp.nn_share_offset(l)), pstride, st), "dss2_gine_bwd")
            gy = gx
            gx_out = gx
        _lib.check(lib.dss2_reduce_partials(_lib.ptr(part), pstride, self.num_partials, self.grad_size, _lib.ptr(flat_grad), 0, st),
                   "dss2_reduce_partials")
        # shared Linear: add the n_conv shares (rows of 72 floats behind the parameters) into the nn.* slot, fixed order
        share = sp.dim_feat * sp.dim_feat + sp.dim_feat
        _lib.check(lib.dss2_reduce_partials(ctypes.c_void_p(flat_grad.data_ptr() + 4 * sp.nn_share_offset(0)), share, sp.n_conv, share,
                                            ctypes.c_void_p(flat_grad.data_ptr() + 4 * self.table["nn.weight"][0]), 0, st),
                   "dss2_reduce_partials")
        return gx_out


class _GINEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, edge_attr, edge_index, runner, pack, names, *params):
        require_cuda()
        out_device = x.device
        named = dict(zip(names, params))
        xg, xs = stage_rows(x)
        eag, eas = stage_rows(edge_attr)
        graph = resolve_graph(edge_index, xg.size(0))
        with torch.cuda.device(xg.device):
            flat = pack.gather(named)
            bufs = runner.alloc(xg.size(0), xg.device, need_grad=any(ctx.needs_input_grad))   # False under torch.no_grad()
            out = runner.forward(graph, xg, xs, eag, eas, flat, bufs)
        ctx.runner, ctx.pack, ctx.names, ctx.graph, ctx.params = runner, pack, names, graph, params
        ctx.saved = (xg, xs, eag, eas, flat, bufs)
        ctx.x_needs_grad, ctx.x_device = x.requires_grad, x.device
        result = out.clone()
        return result if out_device.type == "cuda" else result.to(out_device)

    @staticmethod
    def backward(ctx, grad_out):
        xg, xs, eag, eas, flat, bufs = ctx.saved
        runner = ctx.runner
        go = grad_out.to(device=xg.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(xg.device):
            flat_grad = torch.empty(runner.grad_size, dtype=torch.float32, device=xg.device)
            gx = runner.backward(ctx.graph, xg, xs, eag, eas, flat, bufs, go, flat_grad, need_gx=ctx.x_needs_grad)
        grads = ctx.pack.scatter_grads(flat_grad, dict(zip(ctx.names, ctx.params)))
        gx_ret = gx.clone().to(ctx.x_device) if ctx.x_needs_grad else None
        return (gx_ret, None, None, None, None, None, *grads)


def gine_apply(runner, pack, named_params, x, edge_index, edge_attr):
    names = tuple(named_params.keys())
    return _GINEFunction.apply(x, edge_attr, edge_index, runner, pack, names, *named_params.values())


def make_machinery(spec):
    return GINERunner(spec), ParamPack(spec)
