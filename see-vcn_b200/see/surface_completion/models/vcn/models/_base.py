"""Shared host logic of the VCN models: parameter containers that are state-dict compatible
with the reference, BatchNorm folding, and the call into the C-ABI forward."""
import ctypes

import torch
import torch.nn as nn

from ...... import _abi

PRECISIONS = {"bf16": 0, "fp32": 1, "bf16_layerwise": 2}


def conv_stack(spec):
    """spec: list of ("conv", cin, cout) | ("bn", c) | ("relu",) | ("leaky",) | ("pool",) in reference order,
    so that nn.Sequential numbering (and hence state-dict keys) equals the reference's."""
    mods = []
    for item in spec:
        kind = item[0]
        if kind == "conv":
            mods.append(nn.Conv1d(item[1], item[2], kernel_size=1))
        elif kind == "lin":
            mods.append(nn.Linear(item[1], item[2]))
        elif kind == "bn":
            mods.append(nn.BatchNorm1d(item[1]))
        elif kind == "relu":
            mods.append(nn.ReLU(inplace=True))
        elif kind == "leaky":
            mods.append(nn.LeakyReLU())
        elif kind == "pool":
            mods.append(nn.AdaptiveMaxPool1d(output_size=1))
        else:
            raise ValueError(kind)
    return nn.Sequential(*mods)


def fold_conv_bn(conv, bn=None):
    """(out,in[,1]) weight + bias with an eval-mode BatchNorm folded in -> fp32 (out,in), (out)."""
    w = conv.weight.detach().float().reshape(conv.weight.shape[0], -1)
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w.contiguous(), b.contiguous()


class Encoder(nn.Module):
    """Parameter container for FeatureEncoder (VCN_VC.py:81-94): mlp_conv1 3->128->256, mlp_conv2 512->512->1024."""

    def __init__(self, dims):
        super().__init__()
        self.mlp_conv1 = conv_stack([("conv", dims[0], dims[1]), ("bn", dims[1]), ("relu",), ("conv", dims[1], dims[2])])
        self.mlp_conv2 = conv_stack([("conv", dims[3], dims[4]), ("bn", dims[4]), ("relu",), ("conv", dims[4], dims[5])])


class VCNBase(nn.Module):
    """Inference-only VCN.  ``forward(in_dict)`` keeps the reference contract; the compute is one
    call into ``seevcn_vcn_forward``.  ``precision``: 'bf16' (tcgen05, default) or 'fp32' (SIMT)."""

    viewer_centred = True
    number_coarse = 1024

    def __init__(self, config=None, precision="bf16"):
        super().__init__()
        self.precision = precision
        self._handle = None
        self._handle_key = None
        self._ws = None
        self._plist = None

    # -- packing ---------------------------------------------------------------------
    def _param_key(self):
        # (storage address, in-place version) of every parameter/buffer: changes on load_state_dict, .to(), training
        # steps.  The tensor list itself is cached (walking the module tree costs ~0.3 ms per forward); _apply resets it.
        if self._plist is None:
            self._plist = list(self.parameters()) + list(self.buffers())
        return tuple((p.data_ptr(), p._version) for p in self._plist)

    def _apply(self, fn, *args, **kwargs):
        self._plist = None
        return super()._apply(fn, *args, **kwargs)

    def _folded(self):
        raise NotImplementedError

    def _pack(self):
        key = self._param_key()
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._free()
        if self.training:
            raise RuntimeError("seevcn_b200 VCN is inference-only: call .eval() (BatchNorm is folded)")
        folded = self._folded()
        params = _abi.VcnParams()
        keep = []
        for name, (w, b) in folded.items():
            _abi.require_cuda(w, b)
            keep += [w, b]
            setattr(params, name + "_w", w.data_ptr())
            setattr(params, name + "_b", b.data_ptr())
        params.num_coarse = self.number_coarse
        params.viewer_centred = 1 if self.viewer_centred else 0
        handle = ctypes.c_void_p()
        _abi.check(_abi.lib().seevcn_vcn_create(ctypes.byref(params), ctypes.byref(handle), _abi.stream()))
        torch.cuda.current_stream().synchronize()
        del keep
        self._handle, self._handle_key = handle, key
        return handle

    def _free(self):
        if self._handle is not None:
            _abi.lib().seevcn_vcn_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # -- forward ---------------------------------------------------------------------
    def _run(self, pts, gt_boxes):
        pts = pts.contiguous().float()
        _abi.require_cuda(pts)
        B, N, _ = pts.shape
        dev = pts.device
        with _abi.device_guard(dev):
            handle = self._pack()
            L = _abi.lib()
            ws_bytes = L.seevcn_vcn_workspace_bytes(handle, B, N)
            ws = _abi.workspace(dev, ws_bytes, "vcn")     # per (device, stream): two batches may be in flight on two streams
            coarse = torch.empty((B, self.number_coarse, 3), dtype=torch.float32, device=dev)
            reg_rot = torch.empty((B, 3, 3), dtype=torch.float32, device=dev) if self.viewer_centred else None
            reg_centre = torch.empty((B, 3), dtype=torch.float32, device=dev) if self.viewer_centred else None
            gt = None
            if gt_boxes is not None:
                gt = gt_boxes[:, :7].contiguous().float()
                _abi.require_cuda(gt)
            _abi.check(L.seevcn_vcn_forward(handle, B, N, _abi.ptr(pts), _abi.ptr(gt), _abi.ptr(coarse),
                                            _abi.ptr(reg_rot), _abi.ptr(reg_centre), _abi.ptr(ws),
                                            ws.numel(), PRECISIONS[self.precision], _abi.stream()))
        return coarse, reg_rot, reg_centre
