"""FXencoder and the TCN-based MixFXcloner, same surface as the reference's `networks/architectures.py`.

Mirrors (reference paths relative to /root/reference/mixing_style_transfer/networks/):
  FXencoder  architectures.py:26-70     12 Res_ConvBlocks -> AdaptiveAvgPool1d(1).squeeze(-1)
  TCNModel   architectures.py:76-155    14 FiLM-conditioned dilated blocks -> Conv1d(128->2,k=1) -> clamp(-1,1)
  TCNBlock   architectures.py:177-234

Constructor signatures, attribute names and `state_dict()` key layout are the reference's (so `FXencoder_ps.pt` /
`MixFXcloner_ps.pt` load through inference/style_transfer.py:94-108 unchanged).  `forward` packs the weights once
(BN folding, tap-major split-bf16 tiles) and runs the hand-written sm_100a kernels behind the C ABI
(include/mst_b200.h); eval-mode semantics only, no autograd, no CPU path.
"""
import copy
import ctypes

import torch
import torch.nn as nn

from .. import _cabi
from .network_utils import Conv1d_layer, FiLM, Res_ConvBlock, _param_signature, _Workspace  # noqa: F401


# FXencoder that extracts audio effects from music recordings trained with a contrastive objective
class FXencoder(nn.Module):
    def __init__(self, config):
        super(FXencoder, self).__init__()
        # The reference inserts the stereo input channel into the caller's dict in place (architectures.py:30), which
        # breaks a second construction from the same dict; we work on a copy.
        config = copy.deepcopy(config)
        # input is stereo channeled audio
        config["channels"].insert(0, 2)
        if config["conv_block"] != 'res':
            raise NotImplementedError("FXencoder: only conv_block='res' (inference/configs.yaml:14) has a B200 path")
        self.config = config

        encoder = []
        for i in range(len(config["kernels"])):
            encoder.append(Res_ConvBlock(dimension=1,
                                         in_channels=config["channels"][i],
                                         out_channels=config["channels"][i + 1],
                                         kernel_size=config["kernels"][i],
                                         stride=config["strides"][i],
                                         padding="SAME",
                                         dilation=config["dilation"][i],
                                         norm=config["norm"],
                                         activation=config["activation"],
                                         last_activation=config["activation"]))
        self.encoder = nn.Sequential(*encoder)

        # pooling method
        self.glob_pool = nn.AdaptiveAvgPool1d(1)

        n = len(config["kernels"])
        if n > _cabi.MST_MAX_ENC_BLOCKS:
            raise NotImplementedError(f"FXencoder: at most {_cabi.MST_MAX_ENC_BLOCKS} blocks")
        self._cfg = _cabi.EncConfig()
        self._cfg.n_blocks = n
        for i in range(n + 1):
            self._cfg.channels[i] = config["channels"][i]
        for i in range(n):
            self._cfg.kernels[i] = config["kernels"][i]
            self._cfg.strides[i] = config["strides"][i]
        self._packed = None
        self._packed_sig = None
        self._ws = _Workspace()

    def _pack(self, device):
        sig = _param_signature(self)
        if self._packed is None or sig != self._packed_sig or self._packed.device != device:
            lib = _cabi.lib()
            raw = []
            keep = []
            for blk in self.encoder:
                for conv in (blk.conv1, blk.conv2):
                    for t in conv.raw_pointers():
                        if t is None:
                            raise NotImplementedError("FXencoder: bias=False is not supported by the packed path")
                        t = _cabi.require_cuda_f32(t.detach(), "encoder parameter")
                        keep.append(t)
                        raw.append(t.data_ptr())
            arr = (ctypes.c_void_p * len(raw))(*raw)
            nbytes = lib.mst_enc_packed_bytes(ctypes.byref(self._cfg))
            if nbytes == 0:
                raise RuntimeError("libmst_b200 enc_packed_bytes: " + _cabi.last_error())
            packed = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            off = (-packed.data_ptr()) % 1024
            packed = packed[off:off + nbytes]
            _cabi.check(lib.mst_enc_pack(ctypes.byref(self._cfg), arr, _cabi.ptr(packed), _cabi.current_stream()), "enc_pack")
            torch.cuda.current_stream().synchronize()  # `keep` temporaries may be freed after this returns
            self._packed, self._packed_sig = packed, sig
        return self._packed

    def embed_mean(self, input, scale=None):
        """forward(input) reduced over the batch in the same stream: `scale * sum_b emb[b]`, scale = 1/B by default -- the mean
        embedding of a batch of reference segments (inference/style_transfer.py:152-153); scale=1.0 gives a partial sum."""
        emb = self.forward(input)
        out = torch.empty(emb.shape[1], dtype=torch.float32, device=emb.device)
        _cabi.check(_cabi.lib().mst_rows_reduce(_cabi.ptr(emb), emb.shape[0], emb.shape[1],
                                                float(1.0 / emb.shape[0] if scale is None else scale), _cabi.ptr(out),
                                                _cabi.current_stream()), "rows_reduce")
        return out

    # network forward operation
    def forward(self, input):
        x = _cabi.require_cuda_f32(input, "FXencoder input")
        if x.dim() != 3 or x.shape[1] != 2:
            raise RuntimeError(f"FXencoder expects [B, 2, T], got {tuple(x.shape)}")
        B, _, L = x.shape
        lib = _cabi.lib()
        packed = self._pack(x.device)
        ws_bytes = lib.mst_enc_workspace_bytes(ctypes.byref(self._cfg), B, L)
        ws = self._ws.get(ws_bytes, x.device)
        emb = torch.empty(B, self.config["channels"][-1], dtype=torch.float32, device=x.device)
        _cabi.check(lib.mst_enc_forward(ctypes.byref(self._cfg), _cabi.ptr(packed), _cabi.ptr(x), B, L, _cabi.ptr(emb),
                                        _cabi.ptr(ws), ws.numel(), _cabi.current_stream()), "enc_forward")
        # outputs c feature
        return emb


class _HParams(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


_PRECISIONS = {"f16f8": _cabi.TCN_F16F8, "bf16x3": _cabi.TCN_BF16X3}


def _precision_code(precision: str) -> int:
    """'auto' starts in f16f8 (the caller then checks the range flag); the other names map to the C-ABI constants."""
    if precision == "auto":
        return _cabi.TCN_F16F8
    try:
        return _PRECISIONS[precision]
    except KeyError:
        raise ValueError(f"TCN precision must be 'auto', 'f16f8' or 'bf16x3', got {precision!r}") from None


class _TcnEngine:
    """Packed weights + scratch for one list of TCN blocks (a whole TCNModel, or a stand-alone TCNBlock)."""

    def __init__(self, cfg: "_cabi.TcnConfig"):
        self.cfg = cfg
        self.packed = None
        self.sig = None
        self.ws = _Workspace()

    def pack(self, raw_tensors, sig, device):
        if self.packed is None or sig != self.sig or self.packed.device != device:
            lib = _cabi.lib()
            keep = [_cabi.require_cuda_f32(t.detach(), "TCN parameter") for t in raw_tensors]
            arr = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            nbytes = lib.mst_tcn_packed_bytes(ctypes.byref(self.cfg))
            if nbytes == 0:
                raise NotImplementedError("libmst_b200 tcn config: " + _cabi.last_error())
            packed = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            off = (-packed.data_ptr()) % 1024
            packed = packed[off:off + nbytes]
            _cabi.check(lib.mst_tcn_pack(ctypes.byref(self.cfg), arr, _cabi.ptr(packed), _cabi.current_stream()), "tcn_pack")
            torch.cuda.current_stream().synchronize()  # `keep` temporaries may be freed after this returns
            self.packed, self.sig = packed, sig
        return self.packed

    def workspace(self, B, L, device):
        nbytes = _cabi.lib().mst_tcn_workspace_bytes(ctypes.byref(self.cfg), B, L)
        buf = self.ws.get(nbytes + 1024, device)
        off = (-buf.data_ptr()) % 1024
        return buf[off:off + nbytes]

    def film(self, cond, n_blocks):
        """cond: [n_cond, cond_dim] tensor, or a list with one such tensor per block (SeFa form, architectures.py:139)."""
        lib = _cabi.lib()
        if isinstance(cond, (list, tuple)):
            if len(cond) != n_blocks:
                raise RuntimeError(f"TCN: conditioning list has {len(cond)} entries for {n_blocks} blocks")
            tables = [self.film(c, n_blocks) for c in cond]
            n_cond = {t.shape[1] for t in tables}
            if len(n_cond) != 1:
                raise RuntimeError("TCN: every per-block conditioning tensor must have the same batch size")
            return torch.stack([tables[n][n] for n in range(n_blocks)], dim=0).contiguous()
        c = _cabi.require_cuda_f32(cond, "TCN condition")
        if c.dim() != 2 or c.shape[1] != self.cfg.cond_dim:
            raise RuntimeError(f"TCN condition must be [1 or B, {self.cfg.cond_dim}], got {tuple(c.shape)}")
        out = torch.empty(n_blocks, c.shape[0], self.cfg.channels, 4, dtype=torch.float32, device=c.device)
        _cabi.check(lib.mst_tcn_film_precompute(ctypes.byref(self.cfg), _cabi.ptr(self.packed), _cabi.ptr(c), c.shape[0],
                                                _cabi.ptr(out), _cabi.current_stream()), "tcn_film_precompute")
        return out


# MixFXcloner which is based on a Temporal Convolutional Network (TCN) module
#   original implementation : https://github.com/csteinmetz1/micro-tcn
class TCNModel(nn.Module):
    """ Temporal convolutional network with conditioning module.
        Args:
            nparams (int): Number of conditioning parameters.
            ninputs (int): Number of input channels (mono = 1, stereo 2). Default: 1
            noutputs (int): Number of output channels (mono = 1, stereo 2). Default: 1
            nblocks (int): Number of total TCN blocks. Default: 10
            kernel_size (int): Width of the convolutional kernels. Default: 3
            dialation_growth (int): Compute the dilation factor at each block as dilation_growth ** (n % stack_size). Default: 1
            channel_growth (int): Compute the output channels at each black as in_ch * channel_growth. Default: 2
            channel_width (int): When channel_growth = 1 all blocks use convolutions with this many channels. Default: 64
            stack_size (int): Number of blocks that constitute a single stack of blocks. Default: 10
            grouped (bool): Use grouped convolutions to reduce the total number of parameters. Default: False
            causal (bool): Causal TCN configuration does not consider future input values. Default: False
            skip_connections (bool): Skip connections from each block to the output. Default: False
            num_examples (int): Number of evaluation audio examples to log after each epochs. Default: 4

        Only the configuration the inference path builds (inference/style_transfer.py:48-57: kernel_size 15,
        channel_width 128, channel_growth 1, non-causal, not grouped, 1-2 inputs/outputs) has a B200 path.
        """
    def __init__(self,
                 nparams,
                 ninputs=1,
                 noutputs=1,
                 nblocks=10,
                 kernel_size=3,
                 dilation_growth=1,
                 channel_growth=1,
                 channel_width=32,
                 stack_size=10,
                 cond_dim=2048,
                 grouped=False,
                 causal=False,
                 skip_connections=False,
                 num_examples=4,
                 save_dir=None,
                 **kwargs):
        super(TCNModel, self).__init__()
        self.hparams = _HParams(nparams=nparams, ninputs=ninputs, noutputs=noutputs, nblocks=nblocks,
                                kernel_size=kernel_size, dilation_growth=dilation_growth,
                                channel_growth=channel_growth, channel_width=channel_width, stack_size=stack_size,
                                cond_dim=cond_dim, grouped=grouped, causal=causal, skip_connections=skip_connections,
                                num_examples=num_examples, save_dir=save_dir)
        if channel_growth > 1 or grouped or causal or nparams <= 0:
            raise NotImplementedError("TCNModel: channel_growth>1 / grouped / causal / unconditional variants have no "
                                      "B200 path (the inference entry never builds them)")

        self.blocks = torch.nn.ModuleList()
        for n in range(nblocks):
            in_ch = out_ch if n > 0 else ninputs
            out_ch = self.hparams.channel_width
            dilation = self.hparams.dilation_growth ** (n % self.hparams.stack_size)
            self.blocks.append(TCNBlock(in_ch,
                                        out_ch,
                                        kernel_size=self.hparams.kernel_size,
                                        dilation=dilation,
                                        padding="same" if self.hparams.causal else "valid",
                                        causal=self.hparams.causal,
                                        cond_dim=cond_dim,
                                        grouped=self.hparams.grouped,
                                        conditional=True if self.hparams.nparams > 0 else False))
        self.output = torch.nn.Conv1d(out_ch, noutputs, kernel_size=1)

        cfg = _cabi.TcnConfig(n_blocks=nblocks, n_inputs=ninputs, n_outputs=noutputs, channels=channel_width,
                              kernel_size=kernel_size, dilation_growth=dilation_growth, stack_size=stack_size,
                              cond_dim=cond_dim)
        self._engine = _TcnEngine(cfg)
        # Operand format of the dilated blocks (include/mst_b200.h): "f16f8" (2 tensor units per MMA, activations must
        # stay inside +-448), "bf16x3" (3 units, fp32 range) or "auto" = f16f8 with the range flag read back after the
        # forward and one repeat in bf16x3 when it fired (FiLM's gamma is unbounded, network_utils.py:180-182).
        self.precision = "auto"
        self.last_range_excess = 0.0      # largest |activation| of the last forward that left the f16f8 range (0 = none)
        for n, blk in enumerate(self.blocks):
            blk._bind(self, n)

    def _raw(self):
        raw = []
        for blk in self.blocks:
            raw += blk._raw()
        raw += [self.output.weight, self.output.bias]
        return raw

    def _packed(self, device):
        return self._engine.pack(self._raw(), _param_signature(self), device)

    def forward(self, x, cond, out=None, range_flag=None):
        """x [B, ninputs, T], cond [1 or B, cond_dim] (or one tensor per block) -> [B, noutputs, T].

        Additions to the reference signature (both optional): `out` = preallocated output; `range_flag` = int32[1] CUDA
        tensor that receives the f16f8 range flag WITHOUT a host read-back (for pipelined callers: check it later with
        `flag.view(torch.float32).item() > 0` and call `rerun_bf16x3`)."""
        x = _cabi.require_cuda_f32(x, "TCNModel input")
        hp = self.hparams
        if x.dim() != 3 or x.shape[1] != hp.ninputs:
            raise RuntimeError(f"TCNModel expects [B, {hp.ninputs}, T], got {tuple(x.shape)}")
        B, _, L = x.shape
        eng = self._engine
        eng.pack(self._raw(), _param_signature(self), x.device)
        film = eng.film(cond, hp.nblocks)           # [nblocks, n_cond, 128, 4]
        n_cond = film.shape[1]
        if n_cond not in (1, B):
            raise RuntimeError(f"TCNModel: condition batch {n_cond} does not broadcast against input batch {B}")
        ws = eng.workspace(B, L, x.device)
        out = torch.empty(B, hp.noutputs, L, dtype=torch.float32, device=x.device) if out is None else out
        code = _precision_code(self.precision)
        flag = None
        if code == _cabi.TCN_F16F8 and (self.precision == "auto" or range_flag is not None):
            flag = torch.empty(1, dtype=torch.int32, device=x.device) if range_flag is None else range_flag

        def run(precision_code, flag_t):
            _cabi.check(_cabi.lib().mst_tcn_forward(ctypes.byref(eng.cfg), _cabi.ptr(eng.packed), _cabi.ptr(x),
                                                    _cabi.ptr(film), n_cond, _cabi.ptr(out), B, L, _cabi.ptr(ws),
                                                    ws.numel(), precision_code, _cabi.ptr(flag_t),
                                                    _cabi.current_stream()), "tcn_forward")

        run(code, flag)
        if self.precision == "auto" and range_flag is None:
            # 4-byte read-back (synchronises the stream): did an inter-block activation leave the f16f8 range?
            excess = float(flag.view(torch.float32).item())
            self.last_range_excess = excess
            if excess > 0.0:
                run(_cabi.TCN_BF16X3, None)
        return out

    def rerun_bf16x3(self, x, cond, out):
        """Repeat a forward whose deferred `range_flag` fired (see forward): same arguments, fp32-range operands."""
        saved = self.precision
        self.precision = "bf16x3"
        try:
            return self.forward(x, cond, out=out)
        finally:
            self.precision = saved

    def forward_layers(self, x, cond, on_launch=None):
        """Same computation as forward(), one C-ABI call per kernel launch.  `on_launch(name, phase)` is called with
        phase 'begin' / 'end' around every launch (bench.py records CUDA events there for the roofline leg)."""
        x = _cabi.require_cuda_f32(x, "TCNModel input")
        hp = self.hparams
        B, _, L = x.shape
        eng = self._engine
        eng.pack(self._raw(), _param_signature(self), x.device)
        film = eng.film(cond, hp.nblocks)
        n_cond = film.shape[1]
        ws = eng.workspace(B, L, x.device)
        half = ws.numel() // 2
        act = [ws[:half], ws[half:]]
        out = torch.empty(B, hp.noutputs, L, dtype=torch.float32, device=x.device)
        lib, cfg, st = _cabi.lib(), ctypes.byref(eng.cfg), _cabi.current_stream()
        code = _precision_code(self.precision)
        note = on_launch if on_launch is not None else (lambda name, phase: None)
        note("tcn_block0_kernel", "begin")
        _cabi.check(lib.mst_tcn_block0_forward(cfg, _cabi.ptr(eng.packed), _cabi.ptr(x), _cabi.ptr(film), n_cond,
                                               _cabi.ptr(act[0]), B, L, code, None, st), "tcn_block0_forward")
        note("tcn_block0_kernel", "end")
        cur = 0
        for n in range(1, hp.nblocks):
            last = n == hp.nblocks - 1
            note("tcn_block_umma_kernel", "begin")
            _cabi.check(lib.mst_tcn_layer_forward(cfg, _cabi.ptr(eng.packed), n, _cabi.ptr(act[cur]),
                                                  _cabi.ptr(act[cur ^ 1]), _cabi.ptr(film), n_cond, B, L,
                                                  1 if last else 0, _cabi.ptr(out), code, None, st), "tcn_layer_forward")
            note("tcn_block_umma_kernel", "end")
            cur ^= 1
        return out

    def compute_receptive_field(self):
        """ Compute the receptive field in samples."""
        rf = self.hparams.kernel_size
        for n in range(1, self.hparams.nblocks):
            dilation = self.hparams.dilation_growth ** (n % self.hparams.stack_size)
            rf = rf + ((self.hparams.kernel_size - 1) * dilation)
        return rf


class TCNBlock(torch.nn.Module):
    def __init__(self,
                 in_ch,
                 out_ch,
                 kernel_size=3,
                 dilation=1,
                 cond_dim=2048,
                 grouped=False,
                 causal=False,
                 conditional=False,
                 **kwargs):
        super(TCNBlock, self).__init__()
        if grouped or causal or not conditional:
            raise NotImplementedError("TCNBlock: grouped / causal / unconditional variants have no B200 path")

        self.in_ch = in_ch
        self.out_ch = out_ch
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.grouped = grouped
        self.causal = causal
        self.conditional = conditional

        self.pad_length = ((kernel_size - 1) * dilation) // 2
        self.conv1 = torch.nn.Conv1d(in_ch,
                                     out_ch,
                                     kernel_size=kernel_size,
                                     padding=self.pad_length,
                                     dilation=dilation,
                                     groups=1,
                                     bias=False)
        self.film = FiLM(cond_dim, out_ch)
        self.bn = torch.nn.BatchNorm1d(out_ch)

        self.relu = torch.nn.LeakyReLU()
        self.res = torch.nn.Conv1d(in_ch,
                                   out_ch,
                                   kernel_size=1,
                                   groups=in_ch,
                                   bias=False)
        self._cond_dim = cond_dim
        self.precision = "f16f8"     # stand-alone use only; inside a TCNModel the model's `precision` applies
        # owner = (TCNModel, index) when built by a TCNModel; a stand-alone block gets a private 2-block engine
        self.__dict__["_owner"] = None
        self.__dict__["_own_engine"] = None

    def _bind(self, model, index):
        self.__dict__["_owner"] = (model, index)   # not a registered sub-module: no reference cycle in state_dict

    def _raw(self):
        return [self.conv1.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var,
                self.res.weight, self.film.film_fc.weight, self.film.film_fc.bias]

    def _standalone_engine(self, device):
        """A block used outside a TCNModel: wrap it in a 2-block config whose slot `idx` has this block's dilation."""
        eng = self.__dict__["_own_engine"]
        first = self.in_ch != self.out_ch
        if eng is None:
            cfg = _cabi.TcnConfig(n_blocks=2, n_inputs=self.in_ch if first else 2, n_outputs=2, channels=self.out_ch,
                                  kernel_size=self.kernel_size, dilation_growth=1 if first else self.dilation,
                                  stack_size=2, cond_dim=self._cond_dim)
            if first and self.dilation != 1:
                raise NotImplementedError("stand-alone TCNBlock with in_ch != out_ch must have dilation 1")
            eng = _TcnEngine(cfg)
            self.__dict__["_own_engine"] = eng
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)  # noqa: E731
        ch, k = self.out_ch, self.kernel_size
        dummy_other = [z(ch, ch if first else 2, k), z(ch) + 1, z(ch), z(ch), z(ch) + 1, z(ch, 1, 1),
                       z(2 * ch, self._cond_dim), z(2 * ch)]
        raw = (self._raw() + dummy_other) if first else (dummy_other + self._raw())
        raw += [z(2, ch, 1), z(2)]
        eng.pack(raw, _param_signature(self), device)
        return eng, (0 if first else 1)

    def forward(self, x, p):
        x = _cabi.require_cuda_f32(x, "TCNBlock input")
        if x.dim() != 3 or x.shape[1] != self.in_ch:
            raise RuntimeError(f"TCNBlock expects [B, {self.in_ch}, T], got {tuple(x.shape)}")
        B, _, L = x.shape
        owner = self.__dict__["_owner"]
        if owner is not None:
            model, idx = owner
            eng = model._engine
            eng.pack(model._raw(), _param_signature(model), x.device)
            nblocks = model.hparams.nblocks
            code = _precision_code(model.precision)
        else:
            eng, idx = self._standalone_engine(x.device)
            nblocks = 2
            code = _precision_code(self.precision)
        film = eng.film(p, nblocks)
        n_cond = film.shape[1]
        if n_cond not in (1, B):
            raise RuntimeError(f"TCNBlock: condition batch {n_cond} does not broadcast against input batch {B}")
        ws = eng.workspace(B, L, x.device)
        y = torch.empty(B, self.out_ch, L, dtype=torch.float32, device=x.device)
        _cabi.check(_cabi.lib().mst_tcn_block_forward(ctypes.byref(eng.cfg), _cabi.ptr(eng.packed), idx, _cabi.ptr(x),
                                                      _cabi.ptr(film), n_cond, _cabi.ptr(y), B, L, _cabi.ptr(ws),
                                                      ws.numel(), code, _cabi.current_stream()), "tcn_block_forward")
        return y
