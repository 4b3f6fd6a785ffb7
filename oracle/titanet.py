"""Oracle: TitaNet-L speaker-embedding network.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, module for module, the upstream NeMo classes the reference reaches
through `speaker_embeddings.model_path = "titanet_large"` (helpers.py:281,290):

  nemo/collections/asr/parts/submodules/jasper.py   MaskedConv1d, SqueezeExcite, JasperBlock
  nemo/collections/asr/modules/conv_asr.py          ConvASREncoder, SpeakerDecoder
  nemo/collections/asr/parts/submodules/tdnn_attention.py  TDNNModule, AttentivePoolLayer
  nemo/collections/asr/models/label_models.py       EncDecSpeakerLabelModel.forward
  examples/speaker_tasks/recognition/conf/titanet-large.yaml   (layer table)

SURVEY.md section 8 rows a6, a7.  Weights are random-init under a fixed seed
(BASELINE.json north_star) -- see `seeded_state_dict`.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .features import FilterbankFeatures

# titanet-large.yaml `jasper:` table: (filters, repeat, kernel, residual)
TITANET_L_BLOCKS = [
    (1024, 1, 3, False),
    (1024, 3, 7, True),
    (1024, 3, 11, True),
    (1024, 3, 15, True),
    (3072, 1, 1, False),
]
FEAT_IN = 80
EMB_SIZE = 192
ATTN_CHANNELS = 128
NUM_CLASSES = 16681  # titanet_large checkpoint's classifier width; logits are discarded at inference
SE_REDUCTION = 8
BN_EPS_ENCODER = 1e-3


def get_same_padding(kernel_size, stride=1, dilation=1):
    return (dilation * (kernel_size - 1)) // 2


class MaskedConv1d(nn.Module):
    """jasper.MaskedConv1d with use_mask=True (conv_mask: true): zero the input
    beyond each sequence length, then convolve."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, groups=1, bias=False):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=1, padding=padding, dilation=1, groups=groups, bias=bias)

    def forward(self, x, lens):
        max_len = x.size(2)
        mask = torch.arange(max_len, device=x.device).expand(lens.size(0), max_len) >= lens.unsqueeze(1)
        x = x.masked_fill(mask.unsqueeze(1), 0.0)
        return self.conv(x), lens


class SqueezeExcite(nn.Module):
    """jasper.SqueezeExcite, context_window=-1 (global masked mean), fp32."""

    def __init__(self, channels, reduction_ratio=SE_REDUCTION):
        super().__init__()
        self.fc = nn.Sequential(
            nn.Linear(channels, channels // reduction_ratio, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channels // reduction_ratio, channels, bias=False),
        )

    def forward(self, x, lengths):
        max_len = x.shape[-1]
        valid = torch.arange(max_len, device=x.device).expand(lengths.size(0), max_len) < lengths.unsqueeze(1)
        mask = ~valid.unsqueeze(1)  # 1 = pad
        x = x.float().masked_fill(mask, 0.0)
        y = torch.sum(x, dim=-1, keepdim=True) / (~mask).sum(dim=-1, keepdim=True).type(x.dtype)
        y = y.transpose(1, -1)
        y = self.fc(y)
        y = y.transpose(1, -1)
        y = torch.sigmoid(y)
        return x * y, lengths


class JasperBlock(nn.Module):
    """jasper.JasperBlock, separable=True, normalization='batch', residual_mode='add',
    activation ReLU, dropout inert in eval."""

    def __init__(self, inplanes, planes, repeat, kernel_size, residual, se=True):
        super().__init__()
        padding = get_same_padding(kernel_size)
        conv = nn.ModuleList()
        inplanes_loop = inplanes
        for _ in range(repeat - 1):
            conv.extend(self._conv_bn(inplanes_loop, planes, kernel_size, padding, separable=True))
            conv.append(nn.ReLU())
            inplanes_loop = planes
        conv.extend(self._conv_bn(inplanes_loop, planes, kernel_size, padding, separable=True))
        if se:
            conv.append(SqueezeExcite(planes))
        self.mconv = conv
        if residual:
            self.res = nn.ModuleList([nn.ModuleList(self._conv_bn(inplanes, planes, 1, 0, separable=False))])
        else:
            self.res = None
        self.mout = nn.ReLU()

    @staticmethod
    def _conv_bn(cin, cout, k, padding, separable):
        if separable:
            layers = [
                MaskedConv1d(cin, cin, k, padding=padding, groups=cin, bias=False),
                MaskedConv1d(cin, cout, 1, padding=0, bias=False),
            ]
        else:
            layers = [MaskedConv1d(cin, cout, k, padding=padding, bias=False)]
        layers.append(nn.BatchNorm1d(cout, eps=BN_EPS_ENCODER, momentum=0.1))
        return layers

    def forward(self, xs: List[torch.Tensor], lens_orig):
        out = xs[-1]
        lens = lens_orig
        for layer in self.mconv:
            if isinstance(layer, (MaskedConv1d, SqueezeExcite)):
                out, lens = layer(out, lens)
            else:
                out = layer(out)
        if self.res is not None:
            for i, layer in enumerate(self.res):
                res_out = xs[i]
                for res_layer in layer:
                    if isinstance(res_layer, MaskedConv1d):
                        res_out, _ = res_layer(res_out, lens_orig)
                    else:
                        res_out = res_layer(res_out)
                out = out + res_out
        out = self.mout(out)
        return [out], lens


class ConvASREncoder(nn.Module):
    """conv_asr.ConvASREncoder over the TitaNet-L block table."""

    def __init__(self, feat_in=FEAT_IN, blocks=TITANET_L_BLOCKS):
        super().__init__()
        layers = []
        cin = feat_in
        for filters, repeat, kernel, residual in blocks:
            layers.append(JasperBlock(cin, filters, repeat, kernel, residual))
            cin = filters
        self.encoder = nn.ModuleList(layers)
        self.feat_out = cin

    def forward(self, audio_signal, length):
        xs, lens = [audio_signal], length
        for blk in self.encoder:
            xs, lens = blk(xs, lens)
        return xs[-1], lens


def lens_to_mask(lens, max_len, device=None):
    lens_mat = torch.arange(max_len, device=device)
    mask = lens_mat[:max_len].unsqueeze(0) < lens.unsqueeze(1)
    mask = mask.unsqueeze(1)
    num_values = torch.sum(mask, dim=2, keepdim=True)
    return mask, num_values


def get_statistics_with_mask(x, m, dim=2, eps=1e-10):
    mean = torch.sum((m * x), dim=dim)
    std = torch.sqrt((m * (x - mean.unsqueeze(dim)).pow(2)).sum(dim).clamp(eps))
    return mean, std


class TDNNModule(nn.Module):
    """tdnn_attention.TDNNModule: conv -> ReLU -> BatchNorm."""

    def __init__(self, inp_filters, out_filters, kernel_size=1):
        super().__init__()
        self.conv_layer = nn.Conv1d(inp_filters, out_filters, kernel_size, padding=get_same_padding(kernel_size))
        self.activation = nn.ReLU()
        self.bn = nn.BatchNorm1d(out_filters)

    def forward(self, x):
        return self.bn(self.activation(self.conv_layer(x)))


class AttentivePoolLayer(nn.Module):
    """tdnn_attention.AttentivePoolLayer (channel- and context-dependent statistics pooling)."""

    def __init__(self, inp_filters, attention_channels=ATTN_CHANNELS, eps=1e-10):
        super().__init__()
        self.feat_in = 2 * inp_filters
        self.attention_layer = nn.Sequential(
            TDNNModule(inp_filters * 3, attention_channels, kernel_size=1),
            nn.Tanh(),
            nn.Conv1d(attention_channels, inp_filters, kernel_size=1),
        )
        self.eps = eps

    def forward(self, x, length):
        max_len = x.size(2)
        mask, num_values = lens_to_mask(length, max_len=max_len, device=x.device)
        mean, std = get_statistics_with_mask(x, mask / num_values)
        mean = mean.unsqueeze(2).repeat(1, 1, max_len)
        std = std.unsqueeze(2).repeat(1, 1, max_len)
        attn = torch.cat([x, mean, std], dim=1)
        attn = self.attention_layer(attn)
        attn = attn.masked_fill(mask == 0, -float("inf"))
        alpha = F.softmax(attn, dim=2)
        mu, sg = get_statistics_with_mask(x, alpha)
        return torch.cat((mu, sg), dim=1).unsqueeze(2)


class SpeakerDecoder(nn.Module):
    """conv_asr.SpeakerDecoder(pool_mode='attention', emb_sizes=192, angular=True)."""

    def __init__(self, feat_in, num_classes=NUM_CLASSES, emb_size=EMB_SIZE, compute_logits=True):
        super().__init__()
        self._pooling = AttentivePoolLayer(feat_in)
        self.emb_layers = nn.ModuleList(
            [nn.Sequential(nn.BatchNorm1d(self._pooling.feat_in), nn.Conv1d(self._pooling.feat_in, emb_size, kernel_size=1))]
        )
        self.final = nn.Linear(emb_size, num_classes, bias=False)
        self.compute_logits = compute_logits

    def forward(self, encoder_output, length):
        pool = self._pooling(encoder_output, length)
        emb = None
        for layer in self.emb_layers:
            pool, emb = layer(pool), layer[:2](pool)
        pool = pool.squeeze(-1)
        out = None
        if self.compute_logits:
            pool = F.normalize(pool, p=2, dim=1)
            out = self.final(pool)  # computed and discarded at inference, as upstream
        return out, emb.squeeze(-1)


class TitaNetL(nn.Module):
    """label_models.EncDecSpeakerLabelModel.forward: preprocessor -> encoder -> decoder."""

    def __init__(self, compute_logits=True):
        super().__init__()
        self.preprocessor = FilterbankFeatures()
        self.encoder = ConvASREncoder()
        self.decoder = SpeakerDecoder(self.encoder.feat_out, compute_logits=compute_logits)

    @torch.no_grad()
    def forward(self, input_signal, input_signal_length):
        processed_signal, processed_signal_len = self.preprocessor(input_signal, input_signal_length)
        encoded, length = self.encoder(processed_signal, processed_signal_len)
        logits, embs = self.decoder(encoded, length)
        return logits, embs


def _calibration_audio(gen, n=24, samples=24000):
    """A few seconds of coloured noise + harmonics used only to give the BatchNorm
    running statistics realistic (non-identity) values."""
    t = torch.arange(samples, dtype=torch.float32) / 16000.0
    out = []
    for i in range(n):
        f0 = 90.0 + 160.0 * torch.rand((), generator=gen).item()
        sig = torch.zeros(samples)
        for h in range(1, 25):
            amp = 1.0 / h * (0.3 + torch.rand((), generator=gen).item())
            sig = sig + amp * torch.sin(2 * torch.pi * f0 * h * t + 6.28 * torch.rand((), generator=gen).item())
        sig = sig / sig.abs().max() * 0.3 + 0.01 * torch.randn(samples, generator=gen)
        out.append(sig)
    return torch.stack(out)


def seeded_state_dict(seed: int = 1234, compute_logits: bool = True) -> "TitaNetL":
    """Random-init, fixed-seed TitaNet-L (BASELINE.json north_star).

    Convs/linears: xavier_uniform (jasper.init_weights default).  BatchNorm affine
    parameters are drawn at random and the running statistics are then *calibrated*
    by one training-mode pass over synthetic audio, so that BN is not an identity
    and activations stay O(1) through 14 sub-blocks, as in a trained checkpoint.
    Oracle and CUDA path share this one state_dict, which makes the init
    distribution itself irrelevant to parity."""
    gen = torch.Generator().manual_seed(seed)
    model = TitaNetL(compute_logits=compute_logits)
    for m in model.modules():
        if isinstance(m, (nn.Conv1d, nn.Linear)):
            fan_in, fan_out = nn.init._calculate_fan_in_and_fan_out(m.weight)
            bound = (6.0 / (fan_in + fan_out)) ** 0.5
            with torch.no_grad():
                m.weight.copy_((torch.rand(m.weight.shape, generator=gen) * 2 - 1) * bound)
                if m.bias is not None:
                    m.bias.copy_((torch.rand(m.bias.shape, generator=gen) * 2 - 1) * 0.1)
        elif isinstance(m, nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=gen))
    # calibration pass: BN layers in train mode with momentum 1 record batch statistics
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm1d)]
    model.eval()
    for bn in bns:
        bn.train()
        bn.momentum = 1.0
    audio = _calibration_audio(gen)
    lens = torch.full((audio.shape[0],), audio.shape[1], dtype=torch.long)
    keep = model.decoder.compute_logits
    model.decoder.compute_logits = False
    with torch.no_grad():
        model(audio, lens)
    model.decoder.compute_logits = keep
    for bn in bns:
        bn.eval()
        bn.momentum = 0.1
    model.eval()
    return model
