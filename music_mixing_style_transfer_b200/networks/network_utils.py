"""Building blocks of the networks, same surface as the reference's `networks/network_utils.py`.

Mirrors (reference paths relative to /root/reference/mixing_style_transfer/networks/):
  Conv1d_layer   network_utils.py:15-89     ReflectionPad1d -> Conv1d -> BatchNorm1d -> ReLU
  Res_ConvBlock  network_utils.py:96-119    conv1(x) + x ; conv2(.)
  FiLM           network_utils.py:156-182   Linear(cond) -> split -> r*feature + b

The constructors create exactly the parameter containers the reference does, so `state_dict()` keys and shapes are
identical and the public checkpoints load unchanged.  `forward` does NOT run those torch modules: it calls the sm_100a
kernels in libmst_b200.so (eval-mode semantics; BatchNorm always uses running statistics, as the inference path calls
`.eval()` every iteration, inference/style_transfer.py:148,160).  No CPU path.
"""
import torch
import torch.nn as nn

from .. import _cabi


def _param_signature(module: nn.Module):
    """Cheap change detector for cached packed weights: (ptr, version) of every parameter / buffer."""
    sig = []
    for t in list(module.parameters()) + list(module.buffers()):
        sig.append((t.data_ptr(), t._version))
    return tuple(sig)


class _Workspace:
    """Grow-only device scratch buffers owned by a module (torch-allocated; the library never allocates).

    One buffer per (device, CUDA stream): two forwards of the same module issued on different streams must not share
    their activation ping-pong buffers (the FX path keys its workspace the same way)."""

    def __init__(self):
        self.bufs = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            self.bufs.pop(key, None)     # release the old buffer before growing
            buf = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=device)
            self.bufs[key] = buf
        return buf


# 1-dimensional convolutional layer, in the order of conv -> norm -> activation
class Conv1d_layer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size,
                 stride=1,
                 padding="SAME", dilation=1, bias=True,
                 norm="batch", activation="relu",
                 mode="conv"):
        super(Conv1d_layer, self).__init__()
        if mode != "conv" or padding != "SAME" or dilation != 1 or norm != "batch" or activation != "relu":
            # the FXencoder config (inference/configs.yaml:7-15) only ever builds this combination
            raise NotImplementedError(
                "Conv1d_layer: only mode='conv', padding='SAME', dilation=1, norm='batch', activation='relu' has a "
                f"B200 path (got mode={mode!r}, padding={padding!r}, dilation={dilation}, norm={norm!r}, "
                f"activation={activation!r})")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        pad = int((kernel_size - 1) * dilation)
        self.padding_area = (pad // 2, pad - pad // 2)          # network_utils.py:31-34

        # same container / names as the reference => same state_dict keys (conv1d.conv1d.weight, conv1d.batch_norm.*)
        self.conv1d = nn.Sequential()
        self.conv1d.add_module(f"{mode}1d_pad", nn.ReflectionPad1d(self.padding_area))
        self.conv1d.add_module(f"{mode}1d", nn.Conv1d(in_channels, out_channels, kernel_size,
                                                      stride=stride, padding=0, dilation=dilation, bias=bias))
        self.conv1d.add_module("batch_norm", nn.BatchNorm1d(out_channels))
        self.conv1d.add_module("relu", nn.ReLU())
        self._folded = None
        self._folded_sig = None

    def raw_pointers(self):
        """[w, b, bn_w, bn_b, bn_mean, bn_var] device pointers (b may be NULL)."""
        conv, bn = self.conv1d.conv1d, self.conv1d.batch_norm
        return [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]

    def folded(self):
        """BN-folded, transposed weights w[ci][k][co] and bias (mst_conv1d_fold_bn), cached until a parameter changes."""
        sig = _param_signature(self)
        if self._folded is None or sig != self._folded_sig:
            conv, bn = self.conv1d.conv1d, self.conv1d.batch_norm
            w = _cabi.require_cuda_f32(conv.weight.detach(), "conv weight")
            wf = torch.empty(self.in_channels * self.kernel_size * self.out_channels, dtype=torch.float32, device=w.device)
            bf = torch.empty(self.out_channels, dtype=torch.float32, device=w.device)
            b = None if conv.bias is None else conv.bias.detach().contiguous()
            _cabi.check(_cabi.lib().mst_conv1d_fold_bn(
                _cabi.ptr(w), _cabi.ptr(b), _cabi.ptr(bn.weight.detach()), _cabi.ptr(bn.bias.detach()),
                _cabi.ptr(bn.running_mean), _cabi.ptr(bn.running_var), float(bn.eps),
                self.out_channels, self.in_channels, self.kernel_size, _cabi.ptr(wf), _cabi.ptr(bf),
                _cabi.current_stream()), "conv1d_fold_bn")
            self._folded, self._folded_sig = (wf, bf), sig
        return self._folded

    def forward(self, input, residual=None):
        # input shape should be : batch x channel x time
        x = _cabi.require_cuda_f32(input, "Conv1d_layer input")
        if x.dim() != 3 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"Conv1d_layer expects [B, {self.in_channels}, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        if T <= max(self.padding_area):
            # same condition under which torch's ReflectionPad1d raises
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding "
                               f"{self.padding_area} at dimension 2 of input {list(x.shape)}")
        wf, bf = self.folded()
        t_out = (T + self.stride - 1) // self.stride
        y = torch.empty(B, self.out_channels, t_out, dtype=torch.float32, device=x.device)
        res = None if residual is None else _cabi.require_cuda_f32(residual, "residual")
        _cabi.check(_cabi.lib().mst_enc_conv1d(
            _cabi.ptr(x), _cabi.ptr(wf), _cabi.ptr(bf), _cabi.ptr(res), _cabi.ptr(y),
            B, self.in_channels, T, self.out_channels, self.kernel_size, self.stride, 1, _cabi.current_stream()),
            "enc_conv1d")
        return y


# Residual Block
#   the input is added after the first convolutional layer, retaining its original channel size
#   therefore, the second convolutional layer's output channel may differ
class Res_ConvBlock(nn.Module):
    def __init__(self, dimension,
                 in_channels, out_channels,
                 kernel_size,
                 stride=1, padding="SAME",
                 dilation=1,
                 bias=True,
                 norm="batch",
                 activation="relu", last_activation="relu",
                 mode="conv"):
        super(Res_ConvBlock, self).__init__()
        if dimension != 1:
            raise NotImplementedError("Res_ConvBlock: only dimension=1 exists (the reference's Conv2d_layer is undefined)")
        self.conv1 = Conv1d_layer(in_channels, in_channels, kernel_size, padding=padding, dilation=dilation, bias=bias,
                                  norm=norm, activation=activation)
        self.conv2 = Conv1d_layer(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                  dilation=dilation, bias=bias, norm=norm, activation=last_activation, mode=mode)

    def forward(self, input):
        c1_out = self.conv1(input, residual=input)      # conv1(input) + input, fused (network_utils.py:117)
        c2_out = self.conv2(c1_out)
        return c2_out


# Feature-wise Linear Modulation
class FiLM(nn.Module):
    def __init__(self, condition_len=2048, feature_len=1024):
        super(FiLM, self).__init__()
        self.film_fc = nn.Linear(condition_len, feature_len * 2)
        self.feat_len = feature_len

    def forward(self, feature, condition, sefa=None):
        # Only used stand-alone by callers outside the hot path; inside TCNBlock/TCNModel the modulation is fused into
        # the convolution epilogue (csrc/tcn.cu).  The SeFa branch of the reference relies on torch.eig (removed).
        if sefa:
            raise NotImplementedError("FiLM: the SeFa branch (network_utils.py:163-178) is not part of the inference path")
        film_factor = self.film_fc(condition).unsqueeze(-1)
        r, b = torch.split(film_factor, self.feat_len, dim=1)
        return r * feature + b
