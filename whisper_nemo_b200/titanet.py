"""TitaNet-L speaker-embedding forward on B200: host orchestration over libb200d.so.

Mirrors the operator interface of upstream NeMo's `EncDecSpeakerLabelModel.forward`
(nemo/collections/asr/models/label_models.py: preprocessor -> ConvASREncoder ->
SpeakerDecoder, reached from the reference through helpers.py:281,290
`speaker_embeddings.model_path = "titanet_large"`), but takes segment descriptors
instead of padded audio batches: the waveform stays resident in HBM and the
featurizer kernel gathers / tiles each segment itself.

Weight packing (one-time): BatchNorm folded into the fp16 pointwise weights, the
k=1 depthwise of the last block folded into its pointwise, the TDNN context term
[mean | std] hoisted out of the per-frame GEMM, embedding BN folded into the
final 6144->192 projection, logits layer dropped (inference discards it).
"""
import os
from ctypes import byref
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import GemmEpilogue, ptr

FEAT = 80
FEAT_PAD = 128
EMB = 192
EMB_PAD = 256
HOP = 160
WIN = 400
NFFT = 512

# FilterbankFeatures details that differ between NeMo releases: the default is the classic
# behaviour -- torch.stft's reflect padding and a frame count with "+ 1"; B200D_STFT_PAD_MODE=constant and
# B200D_SEQ_LEN_PLUS_ONE=0 select what later releases (and transformers' port of the featurizer) do.
FEAT_ZERO_PAD, FEAT_NO_PLUS_ONE = 1, 2  # include/b200d.h
FEATURIZER_VARIANT = ((FEAT_ZERO_PAD if os.environ.get("B200D_STFT_PAD_MODE", "reflect") == "constant" else 0)
                      | (FEAT_NO_PLUS_ONE if os.environ.get("B200D_SEQ_LEN_PLUS_ONE", "1") == "0" else 0))


def frames_of(fixed_len: int) -> int:
    """Feature frames of a window of `fixed_len` samples (FilterbankFeatures.get_seq_len at n_fft 512, hop 160)."""
    return fixed_len // HOP + (0 if FEATURIZER_VARIANT & FEAT_NO_PLUS_ONE else 1)


# ------------------------------------------------------------------------------------------ tables
def slaney_mel_filterbank(sr: int = 16000, n_fft: int = NFFT, n_mels: int = FEAT) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular filterbank (what librosa.filters.mel returns with
    its defaults, which is what NeMo's FilterbankFeatures uses).  float32 [n_mels, n_fft//2+1]."""
    f_sp, brk = 200.0 / 3.0, 1000.0
    brk_mel, step = brk / f_sp, np.log(6.4) / 27.0

    def to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= brk, brk_mel + np.log(np.maximum(f, 1e-12) / brk) / step, f / f_sp)

    def to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= brk_mel, brk * np.exp(step * (m - brk_mel)), f_sp * m)

    edges = to_hz(np.linspace(to_mel(0.0), to_mel(sr / 2.0), n_mels + 2))
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    up = (bins[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    down = (edges[2:, None] - bins[None, :]) / (edges[2:] - edges[1:-1])[:, None]
    fb = np.clip(np.minimum(up, down), 0.0, None)
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb.astype(np.float32)


def pack_filterbank(fb: np.ndarray):
    """Sparse (start bin, offset, packed weights) form consumed by b200d_featurize_windows / b200d_mel_stream."""
    starts, offs, weights = [], [0], []
    for m in range(fb.shape[0]):
        nz = np.nonzero(fb[m])[0]
        if len(nz) == 0:
            starts.append(0)
            offs.append(offs[-1])
            continue
        a, b = int(nz[0]), int(nz[-1]) + 1
        starts.append(a)
        weights.extend(fb[m, a:b].tolist())
        offs.append(offs[-1] + (b - a))
    return np.asarray(starts, np.int32), np.asarray(offs, np.int32), np.asarray(weights, np.float32)


# ------------------------------------------------------------------------------------------ weights
@dataclass
class SubBlock:
    ksize: int
    dw: Optional[torch.Tensor]  # float32 [k][Cin_pad] or None (k == 1 folded)
    w: torch.Tensor  # fp16 [Cout][Cin_pad]
    bias: torch.Tensor  # float32 [Cout]


@dataclass
class Block:
    subs: List[SubBlock]
    se_w1: torch.Tensor  # fp16 [C/8][C]
    se_w2: torch.Tensor  # fp16 [C][C/8]
    res_w: Optional[torch.Tensor]  # fp16 [Cout][Cin]
    res_bias: Optional[torch.Tensor]
    cout: int


@dataclass
class PackedTitaNet:
    blocks: List[Block] = field(default_factory=list)
    tdnn_wx: torch.Tensor = None  # fp16 [128][3072]
    tdnn_wctx: torch.Tensor = None  # fp16 [128][6144]
    tdnn_b: torch.Tensor = None  # float32 [128]
    tdnn_scale: torch.Tensor = None
    tdnn_shift: torch.Tensor = None
    attn_w2: torch.Tensor = None  # fp16 [3072][128]
    attn_b2: torch.Tensor = None
    emb_w: torch.Tensor = None  # fp16 [256][6144] (rows >= 192 zero)
    emb_b: torch.Tensor = None  # float32 [256]
    zeros: torch.Tensor = None  # float32 [3072] zero bias
    fb_start: torch.Tensor = None
    fb_off: torch.Tensor = None
    fb_w: torch.Tensor = None
    window: torch.Tensor = None


def _bn_fold(sd, prefix, eps):
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    m, v = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    s = g / torch.sqrt(v + eps)
    return s, b - m * s


def pack_weights(state_dict: Dict[str, torch.Tensor], device="cuda") -> PackedTitaNet:
    """state_dict with upstream NeMo's TitaNet-L key layout (`encoder.encoder.{i}.mconv.{j}...`,
    `encoder.encoder.{i}.res.0.{0,1}...`, `decoder._pooling.attention_layer...`,
    `decoder.emb_layers.0.{0,1}...`).  Sub-module indices are discovered from the keys, so both
    NeMo's [conv, conv, bn, act, dropout] and a dropout-free layout load."""
    sd = {k: v.detach().cpu() for k, v in state_dict.items()}
    pk = PackedTitaNet()
    n_blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.encoder."))
    f16 = lambda t: t.to(torch.float16).contiguous().to(device)
    f32 = lambda t: t.to(torch.float32).contiguous().to(device)
    for i in range(n_blocks):
        pre = f"encoder.encoder.{i}.mconv."
        idx = sorted({int(k[len(pre):].split(".")[0]) for k in sd if k.startswith(pre)})
        convs = [j for j in idx if f"{pre}{j}.conv.weight" in sd]
        bns = [j for j in idx if f"{pre}{j}.running_mean" in sd]
        ses = [j for j in idx if f"{pre}{j}.fc.0.weight" in sd]
        assert len(convs) == 2 * len(bns) and len(ses) == 1, f"unexpected layout in block {i}"
        subs = []
        for r, bn_j in enumerate(bns):
            dw_w = sd[f"{pre}{convs[2 * r]}.conv.weight"].double()  # [Cin,1,k]
            pw_w = sd[f"{pre}{convs[2 * r + 1]}.conv.weight"].double()[:, :, 0]  # [Cout,Cin]
            s, t = _bn_fold(sd, f"{pre}{bn_j}", 1e-3)
            cin, k = dw_w.shape[0], dw_w.shape[2]
            cin_pad = FEAT_PAD if cin == FEAT else cin
            w = pw_w * s[:, None]
            dw = None
            if k == 1:
                w = w * dw_w[:, 0, 0][None, :]
            else:
                dwp = torch.zeros(k, cin_pad, dtype=torch.float64)
                dwp[:, :cin] = dw_w[:, 0, :].t()
                dw = f32(dwp)
            wp = torch.zeros(w.shape[0], cin_pad, dtype=torch.float64)
            wp[:, :cin] = w
            subs.append(SubBlock(ksize=k, dw=dw, w=f16(wp), bias=f32(t)))
        se_pre = f"{pre}{ses[0]}.fc."
        res_w = res_b = None
        rpre = f"encoder.encoder.{i}.res.0."
        if f"{rpre}0.conv.weight" in sd:
            rw = sd[f"{rpre}0.conv.weight"].double()[:, :, 0]
            s, t = _bn_fold(sd, f"{rpre}1", 1e-3)
            res_w, res_b = f16(rw * s[:, None]), f32(t)
        pk.blocks.append(
            Block(subs=subs, se_w1=f16(sd[se_pre + "0.weight"]), se_w2=f16(sd[se_pre + "2.weight"]), res_w=res_w, res_bias=res_b,
                  cout=subs[-1].w.shape[0])
        )
    c_enc = pk.blocks[-1].cout
    ap = "decoder._pooling.attention_layer."
    tw = sd[ap + "0.conv_layer.weight"].double()[:, :, 0]  # [128, 3*C]
    pk.tdnn_wx = f16(tw[:, :c_enc])
    pk.tdnn_wctx = f16(tw[:, c_enc:])
    pk.tdnn_b = f32(sd[ap + "0.conv_layer.bias"])
    s, t = _bn_fold(sd, ap + "0.bn", 1e-5)
    pk.tdnn_scale, pk.tdnn_shift = f32(s), f32(t)
    pk.attn_w2 = f16(sd[ap + "2.weight"][:, :, 0])
    pk.attn_b2 = f32(sd[ap + "2.bias"])
    s, t = _bn_fold(sd, "decoder.emb_layers.0.0", 1e-5)
    ew = sd["decoder.emb_layers.0.1.weight"].double()[:, :, 0]  # [192, 6144]
    eb = sd["decoder.emb_layers.0.1.bias"].double() + ew @ t
    ewp = torch.zeros(EMB_PAD, ew.shape[1], dtype=torch.float64)
    ewp[: ew.shape[0]] = ew * s[None, :]
    ebp = torch.zeros(EMB_PAD, dtype=torch.float64)
    ebp[: eb.shape[0]] = eb
    pk.emb_w, pk.emb_b = f16(ewp), f32(ebp)
    pk.zeros = torch.zeros(4096, dtype=torch.float32, device=device)
    fs, fo, fw = pack_filterbank(slaney_mel_filterbank())
    pk.fb_start = torch.from_numpy(fs).to(device)
    pk.fb_off = torch.from_numpy(fo).to(device)
    pk.fb_w = torch.from_numpy(fw).to(device)
    pk.window = torch.hann_window(WIN, periodic=False, dtype=torch.float64).float().to(device)
    return pk


# ------------------------------------------------------------------------------------------ kernels
def gemm(A, W, out, mode, M=None, bias=None, scale=None, shift=None, rowvec=None, aux16=None, rows_per_seg=0, colsum=None):
    """out = epilogue(A[M,K] @ W[N,K]^T) through b200d_gemm_f16."""
    M = A.shape[0] if M is None else M
    N, K = W.shape
    epi = GemmEpilogue()
    epi.mode = mode
    epi.rows_per_seg = rows_per_seg
    epi.bias = bias.data_ptr() if bias is not None else None
    epi.scale = scale.data_ptr() if scale is not None else None
    epi.shift = shift.data_ptr() if shift is not None else None
    epi.rowvec = rowvec.data_ptr() if rowvec is not None else None
    epi.aux16 = aux16.data_ptr() if aux16 is not None else None
    epi.flags = _cabi.gemm_flags()
    epi.colsum = colsum.data_ptr() if colsum is not None else None
    _cabi.call("b200d_gemm_f16", ptr(A), A.stride(0), ptr(W), W.stride(0), M, N, K, ptr(out), out.stride(0), byref(epi), _cabi._stream())
    return out


def featurize(pk: PackedTitaNet, wav: torch.Tensor, seg_start: torch.Tensor, seg_len: torch.Tensor, fixed_len: int,
              out16: torch.Tensor = None, want_f32: bool = False, logmel: torch.Tensor = None, seg_row0: torch.Tensor = None,
              n_on_stream: int = None):
    """wav float32 [n] on device; seg_start / seg_len int32 [n_seg] on device.  `logmel` / `seg_row0`: the recording's
    stream frames (mel_stream) and each segment's first row in them -- interior frames are then copied, not recomputed;
    `n_on_stream`: how many leading segments are on a stream (the caller orders them first; default: unknown -> 0, every
    segment's non-interior frames are then computed by the spread-out kernel).
    Returns (fp16 [n_seg*T, 128], optional float32 [n_seg, T, 80])."""
    n_seg = seg_start.numel()
    T = frames_of(fixed_len)
    if out16 is None:
        out16 = torch.empty(n_seg * T, FEAT_PAD, dtype=torch.float16, device=wav.device)
    out32 = torch.empty(n_seg, T, FEAT, dtype=torch.float32, device=wav.device) if want_f32 else None
    if logmel is None or seg_row0 is None:
        logmel = seg_row0 = None
    scratch = torch.empty(n_seg * T, FEAT, dtype=torch.float32, device=wav.device)
    _cabi.call("b200d_featurize_windows", ptr(wav), wav.numel(), ptr(logmel), ptr(seg_start), ptr(seg_len), ptr(seg_row0),
               int(n_on_stream or 0) if logmel is not None else 0, n_seg, fixed_len,
               ptr(pk.fb_start), ptr(pk.fb_off), ptr(pk.fb_w), pk.fb_w.numel(), ptr(pk.window), FEATURIZER_VARIANT, ptr(scratch), ptr(out16),
               out16.stride(0), ptr(out32), _cabi._stream())
    return out16, out32


STREAM_PAD = 32  # rows of a stream are padded to the 32 frames one CTA of mel_stream_kernel computes


def plan_mel_streams(start: np.ndarray, length: np.ndarray, fixed: np.ndarray):
    """Host plan of the once-per-recording frames.  start / length / fixed: int64 [n] first sample (in the concatenated
    waveform), true length and tiled-up length of every window of every scale.
    Windows that fill their fixed length and share a grid phase (start mod 160) are chained into STREAMS wherever they
    overlap or touch; frame g of a stream is centred on sample stream_start + 160 g.
    Returns (stream_start int64 [S], stream_off int32 [S + 1] (rows, multiples of 32), row0 int32 [n]: the row of the frame
    centred on each window's first sample, -1 for windows without a stream (tiled ones))."""
    start, length, fixed = (np.asarray(a, dtype=np.int64) for a in (start, length, fixed))
    row0 = np.full(start.shape[0], -1, dtype=np.int64)
    full = np.nonzero(length == fixed)[0]
    s_start, s_rows = [], []
    rows_total = 0
    phase = start[full] % HOP
    for ph in np.unique(phase):
        idx = full[phase == ph]
        idx = idx[np.argsort(start[idx], kind="stable")]
        s, e = start[idx], start[idx] + fixed[idx]
        reach = np.maximum.accumulate(e)
        first = np.concatenate([[True], s[1:] > reach[:-1]])  # a window beyond everything seen so far opens a new stream
        gid = np.cumsum(first) - 1
        a = s[first]
        last = np.concatenate([first[1:], [True]])
        rows = (reach[last] - a) // HOP + 1
        rows = (rows + STREAM_PAD - 1) // STREAM_PAD * STREAM_PAD
        off = rows_total + np.concatenate([[0], np.cumsum(rows)[:-1]])
        row0[idx] = off[gid] + (s - a[gid]) // HOP
        s_start.append(a)
        s_rows.append(rows)
        rows_total += int(rows.sum())
    if not s_start:
        return np.zeros(0, np.int64), np.zeros(1, np.int32), row0.astype(np.int32)
    s_start, s_rows = np.concatenate(s_start), np.concatenate(s_rows)
    stream_off = np.concatenate([[0], np.cumsum(s_rows)])
    if stream_off[-1] >= 2 ** 31:
        raise ValueError("more than 2^31 stream frames; split the batch")
    return s_start.astype(np.int64), stream_off.astype(np.int32), row0.astype(np.int32)


def mel_stream(pk: PackedTitaNet, wav: torch.Tensor, stream_start: torch.Tensor, stream_off: torch.Tensor, total_rows: int, out: torch.Tensor = None):
    """log-mel rows float32 [total_rows, 80] of the planned streams (device arrays from plan_mel_streams)."""
    if out is None:
        out = torch.empty(total_rows, FEAT, dtype=torch.float32, device=wav.device)
    _cabi.call("b200d_mel_stream", ptr(wav), wav.numel(), ptr(stream_start), ptr(stream_off), stream_start.numel(), total_rows,
               ptr(pk.fb_start), ptr(pk.fb_off), ptr(pk.fb_w), pk.fb_w.numel(), ptr(pk.window), ptr(out), _cabi._stream())
    return out


class Workspace:
    """Activation buffers for up to `max_frames` frames / `max_segs` segments, allocated once."""

    def __init__(self, max_frames: int, max_segs: int, device):
        h = lambda r, c: torch.empty(r, c, dtype=torch.float16, device=device)
        self.max_frames, self.max_segs = max_frames, max_segs
        self.x0 = h(max_frames, FEAT_PAD)
        self.a = h(max_frames, 1024)
        self.b = h(max_frames, 1024)
        self.y = h(max_frames, 1024)
        self.d = h(max_frames, 1024)
        self.x = h(max_frames, 3072)
        self.e = h(max_frames, 3072)
        self.hid = h(max_frames, 128)
        self.mean16 = h(max_segs, 3072)
        self.sehid = h(max_segs, 384)
        self.gate = torch.empty(max_segs, 3072, dtype=torch.float32, device=device)
        self.stats16 = h(max_segs, 6144)
        self.segbias = torch.empty(max_segs, 128, dtype=torch.float32, device=device)
        self.pool16 = h(max_segs, 6144)
        self.emb = torch.empty(max_segs, EMB_PAD, dtype=torch.float32, device=device)
        self.colsum = torch.empty((max_frames // 32 + 1) * 2 * 3072, dtype=torch.float32, device=device)


def _flat(buf, rows, cols):
    """Contiguous [rows, cols] view on the front of a larger scratch buffer."""
    return buf.view(-1)[: rows * cols].view(rows, cols)


def _se_gate(pk, ws, blk: Block, y, n_seg, T, C, from_colsum=False):
    mean16 = _flat(ws.mean16, n_seg, C)
    hid = _flat(ws.sehid, n_seg, C // 8)
    gate = _flat(ws.gate, n_seg, C)
    if from_colsum:  # the GEMM that produced y left per-32-row column sums (GemmEpilogue.colsum)
        _cabi.call("b200d_se_mean_from_colsum", ptr(ws.colsum), n_seg, T, C, ptr(mean16), _cabi._stream())
    else:
        _cabi.call("b200d_time_stats", ptr(y), n_seg, T, C, 0, ptr(mean16), _cabi._stream())
    gemm(mean16, blk.se_w1, hid, _cabi.EPI_BIAS_RELU, bias=pk.zeros)
    gemm(hid, blk.se_w2, gate, _cabi.EPI_SIGMOID_F32)
    return gate


def _se_input_gemm(ws, A, W, out, M, T, bias):
    """The pointwise conv in front of a squeeze-excite; when it runs on the CTA-pair kernel its epilogue also leaves the
    column sums the SE time mean needs.  Returns whether it did."""
    cs = T >= 32 and _cabi.gemm_uses_pair(M, W.shape[0], _cabi.EPI_BIAS, _cabi.gemm_flags())
    gemm(A, W, out, _cabi.EPI_BIAS, M=M, bias=bias, rows_per_seg=T if cs else 0, colsum=ws.colsum if cs else None)
    return cs


def forward_frames(pk: PackedTitaNet, ws: Workspace, n_seg: int, T: int, taps: dict = None) -> torch.Tensor:
    """Encoder + decoder over ws.x0[: n_seg*T] (already featurized).  Returns emb float32 [n_seg, 192] (a view)."""
    M = n_seg * T
    s = _cabi._stream()
    dwc = lambda x, y, w, C, k: _cabi.call("b200d_depthwise_conv", ptr(x), ptr(y), ptr(w), n_seg, T, C, k, s)
    # ---- block 0: dw3 -> pw 80->1024 -> BN -> SE -> ReLU
    b0 = pk.blocks[0]
    d0v = _flat(ws.d, M, FEAT_PAD)
    dwc(ws.x0, d0v, b0.subs[0].dw, FEAT_PAD, b0.subs[0].ksize)
    cs = _se_input_gemm(ws, d0v, b0.subs[0].w, ws.y, M, T, b0.subs[0].bias)
    gate = _se_gate(pk, ws, b0, ws.y, n_seg, T, 1024, cs)
    _cabi.call("b200d_se_apply_relu", ptr(ws.y), ptr(gate), ptr(ws.a), n_seg, T, 1024, s)
    cur, nxt = ws.a, ws.b
    if taps is not None:
        taps["block0"] = cur[:M].clone()
    # ---- blocks 1..3: 3 x (dw -> pw -> BN [-> ReLU]) -> SE ; + BN(conv1x1(in)) ; ReLU
    for bi in (1, 2, 3):
        blk = pk.blocks[bi]
        src = cur
        for r, sb in enumerate(blk.subs):
            dwc(src, ws.d, sb.dw, 1024, sb.ksize)
            if r == len(blk.subs) - 1:
                cs = _se_input_gemm(ws, ws.d, sb.w, ws.y, M, T, sb.bias)
            else:
                gemm(ws.d, sb.w, ws.y, _cabi.EPI_BIAS_RELU, M=M, bias=sb.bias)
            src = ws.y  # next depthwise reads y and writes d; its GEMM then overwrites y
        gate = _se_gate(pk, ws, blk, ws.y, n_seg, T, 1024, cs)
        gemm(cur, blk.res_w, nxt, _cabi.EPI_SE_RES, M=M, bias=blk.res_bias, rowvec=gate, aux16=ws.y, rows_per_seg=T)
        cur, nxt = nxt, cur
        if taps is not None:
            taps[f"block{bi}"] = cur[:M].clone()
    # ---- block 4: (dw k=1 folded) pw 1024->3072 -> BN -> SE -> ReLU
    b4 = pk.blocks[4]
    cs = _se_input_gemm(ws, cur, b4.subs[0].w, ws.e, M, T, b4.subs[0].bias)
    gate = _se_gate(pk, ws, b4, ws.e, n_seg, T, 3072, cs)
    stats16 = ws.stats16[:n_seg]
    _cabi.call("b200d_se_apply_relu_stats", ptr(ws.e), ptr(gate), ptr(ws.x), n_seg, T, 3072, ptr(stats16), s)
    if taps is not None:
        taps["encoder"] = ws.x[:M].clone()
    # ---- decoder: attentive statistics pooling + embedding projection
    segbias = ws.segbias[:n_seg]
    gemm(stats16, pk.tdnn_wctx, segbias, _cabi.EPI_BIAS_F32, bias=pk.tdnn_b)
    gemm(ws.x, pk.tdnn_wx, ws.hid, _cabi.EPI_TDNN, M=M, scale=pk.tdnn_scale, shift=pk.tdnn_shift, rowvec=segbias, rows_per_seg=T)
    gemm(ws.hid, pk.attn_w2, ws.e, _cabi.EPI_BIAS, M=M, bias=pk.attn_b2)
    pool16 = ws.pool16[:n_seg]
    _cabi.call("b200d_attn_pool", ptr(ws.x), ptr(ws.e), n_seg, T, 3072, ptr(pool16), s)
    if taps is not None:
        taps["pool"] = pool16.clone()
    emb = ws.emb[:n_seg]
    gemm(pool16, pk.emb_w, emb, _cabi.EPI_BIAS_F32, bias=pk.emb_b)
    return emb[:, :EMB]


def pack_weights_cabi(state_dict: Dict[str, torch.Tensor], device="cuda"):
    """b200d_titanet_pack_weights: the same folding / padding / fp16 conversion as `pack_weights`, done by the library (C++, host)
    into ONE position-independent blob.  Returns (TitaNetDesc, uint8 tensor on `device`)."""
    import ctypes

    sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items()
          if torch.is_tensor(v) and v.dtype.is_floating_point}
    if "preprocessor.featurizer.fb" not in sd:  # NeMo checkpoints carry both buffers; the seeded random init does not
        sd["preprocessor.featurizer.fb"] = torch.from_numpy(slaney_mel_filterbank()).unsqueeze(0).contiguous()
    if "preprocessor.featurizer.window" not in sd:
        sd["preprocessor.featurizer.window"] = torch.hann_window(WIN, periodic=False, dtype=torch.float64).float()
    names = list(sd)
    n = len(names)
    c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
    c_data = (ctypes.c_void_p * n)(*[sd[k].data_ptr() for k in names])
    c_numel = (ctypes.c_int64 * n)(*[sd[k].numel() for k in names])
    desc = _cabi.TitaNetDesc()
    lib = _cabi.load()
    _cabi.check(lib.b200d_titanet_pack_weights(n, c_names, c_data, c_numel, ctypes.byref(desc), None, 0), "b200d_titanet_pack_weights")
    blob = torch.empty(int(desc.packed_bytes), dtype=torch.uint8)
    _cabi.check(lib.b200d_titanet_pack_weights(n, c_names, c_data, c_numel, ctypes.byref(desc), blob.data_ptr(), blob.numel()), "b200d_titanet_pack_weights")
    return desc, blob.to(device)


class TitaNetB200:
    """Speaker-embedding extractor: `embed_segments` is the device-side replacement of the
    `_extract_embeddings` dataloader loop of upstream ClusteringDiarizer -- one `b200d_titanet_forward` call per
    (scale, window length): the ~60 kernels per group of windows are launched from C++."""

    def __init__(self, state_dict, device="cuda", max_frames: int = None):
        _cabi.require_device()
        if max_frames is None:  # frames per launch group: ~21 KB of fp16 activations per frame (2.7 GB at the default)
            max_frames = int(os.environ.get("B200D_MAX_FRAMES", 131072))
        self.device = torch.device(device)
        self._state_dict = state_dict
        self._pk = None
        self.desc, self.packed = pack_weights_cabi(state_dict, self.device)
        self.max_frames = max_frames
        self._ws = {}   # one activation workspace per launching stream (concurrent scales must not share buffers)
        self._pyws = {}

    @property
    def pk(self) -> PackedTitaNet:
        """The Python-side packing (per-operand tensors): only the step-by-step reference orchestration (`forward_frames`,
        `featurize`, taps for the tests) uses it; the product path reads the library's blob."""
        if self._pk is None:
            self._pk = pack_weights(self._state_dict, self.device)
        return self._pk

    def _workspace(self):
        key = torch.cuda.current_stream().cuda_stream
        ws = self._ws.get(key)
        if ws is None:
            import ctypes

            nbytes = _cabi.load().b200d_titanet_workspace_bytes(ctypes.byref(self.desc), self.max_frames, max(1024, self.max_frames // 32))
            ws = self._ws[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return ws

    def group_windows(self, fixed_len: int) -> int:
        """Windows of `fixed_len` samples per launch group of embed_segments (b200d_titanet_group_windows)."""
        import ctypes

        return int(_cabi.load().b200d_titanet_group_windows(ctypes.byref(self.desc), self._workspace().numel(), int(fixed_len), FEATURIZER_VARIANT))

    def mel_stream(self, wav: torch.Tensor, stream_start: torch.Tensor, stream_off: torch.Tensor, total_rows: int) -> torch.Tensor:
        """log-mel rows float32 [total_rows, 80] of the planned streams (plan_mel_streams), each frame computed once."""
        import ctypes

        out = torch.empty(total_rows, FEAT, dtype=torch.float32, device=wav.device)
        _cabi.call("b200d_titanet_mel_stream", ctypes.byref(self.desc), ptr(self.packed), ptr(wav), wav.numel(), ptr(stream_start), ptr(stream_off),
                   stream_start.numel(), total_rows, ptr(out), _cabi._stream())
        return out

    @torch.no_grad()
    def embed_segments(self, wav: torch.Tensor, seg_start: torch.Tensor, seg_len: torch.Tensor, fixed_len: int, taps: dict = None,
                       logmel: torch.Tensor = None, seg_row0: torch.Tensor = None, n_on_stream: int = None):
        """All segments share `fixed_len` (batch max under fixed_seq collate).  `logmel` / `seg_row0` / `n_on_stream`: see
        featurize.  Returns float32 [n_seg, 192].  `taps` (a dict) selects the step-by-step reference orchestration and receives the
        intermediate activations."""
        if taps is not None:
            return self._embed_segments_stepwise(wav, seg_start, seg_len, fixed_len, taps, logmel, seg_row0, n_on_stream)
        import ctypes

        n_seg = seg_start.numel()
        out = torch.empty(n_seg, EMB, dtype=torch.float32, device=self.device)
        if logmel is None or seg_row0 is None:
            logmel = seg_row0 = None
        ws = self._workspace()
        _cabi.call("b200d_titanet_forward", ctypes.byref(self.desc), ptr(self.packed), ptr(wav), wav.numel(), ptr(logmel), ptr(seg_start), ptr(seg_len),
                   ptr(seg_row0), int(n_on_stream or 0) if logmel is not None else 0, n_seg, fixed_len, FEATURIZER_VARIANT, _cabi.gemm_flags(), ptr(out), out.stride(0), ptr(ws), ws.numel(), _cabi._stream())
        return out

    def _py_workspace(self, T):
        key = torch.cuda.current_stream().cuda_stream
        max_segs = max(1, self.max_frames // T)
        ws = self._pyws.get(key)
        if ws is None or ws.max_segs < max_segs:
            ws = self._pyws[key] = Workspace(self.max_frames, max(max_segs, 1024), self.device)
        return ws

    def _embed_segments_stepwise(self, wav, seg_start, seg_len, fixed_len, taps, logmel=None, seg_row0=None, n_on_stream=None):
        n_seg = seg_start.numel()
        T = frames_of(fixed_len)
        ws = self._py_workspace(T)
        segs_per_chunk = max(1, self.max_frames // T)
        out = torch.empty(n_seg, EMB, dtype=torch.float32, device=self.device)
        for c0 in range(0, n_seg, segs_per_chunk):
            c1 = min(n_seg, c0 + segs_per_chunk)
            featurize(self.pk, wav, seg_start[c0:c1], seg_len[c0:c1], fixed_len, out16=ws.x0, logmel=logmel,
                      seg_row0=None if seg_row0 is None else seg_row0[c0:c1], n_on_stream=max(0, min(int(n_on_stream or 0) - c0, c1 - c0)))
            out[c0:c1] = forward_frames(self.pk, ws, c1 - c0, T, taps)
        return out
