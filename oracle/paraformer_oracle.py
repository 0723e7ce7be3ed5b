"""ORACLE — CPU fp32 restatement of FunASR's offline Paraformer hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product (funasr_b200/) never does and fails loudly without its CUDA library.

Parity status: PINNED against the live reference.  oracle/make_golden.py runs the unmodified
reference classes from /root/reference (funasr 1.4.3) on CPU with the seeded synthetic weights of
funasr_b200/synth.py and stores their outputs under tests/golden/; tests/test_oracle_golden.py checks
this restatement against those files, and (when /root/reference is present) against the live
reference directly.  The Fbank arithmetic lives in a third-party dependency that is NOT under
/root/reference: torchaudio.compliance.kaldi.fbank — pinned here to torchaudio 2.11.0+cu128
(site-packages/torchaudio/compliance/kaldi.py); the reference ships no test that pins Fbank values,
so that boundary is pinned by our own golden vectors generated from that torchaudio version.

Each function cites the reference lines it restates (paths relative to /root/reference).
Plain torch CPU ops are used in the same order as the reference so that on the same machine the
restatement is (near) bit-identical to it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
EPS_F32 = torch.finfo(torch.float32).eps  # kaldi.py:_get_epsilon -> 1.1920929e-07


# --------------------------------------------------------------------------------------
# Frontend: torchaudio.compliance.kaldi.fbank (kaldi.py:514-647) as called by
# funasr/frontends/wav_frontend.py:165-189, then apply_lfr (:63-86) and apply_cmvn (:46-60)
# --------------------------------------------------------------------------------------
def mel_scale(f):
    return 1127.0 * (1.0 + f / 700.0).log()          # kaldi.py:mel_scale


def mel_scale_scalar(f: float) -> float:
    return 1127.0 * math.log(1.0 + f / 700.0)        # kaldi.py:mel_scale_scalar


def get_mel_banks(num_bins=80, padded=512, fs=16000.0, low=20.0, high=0.0) -> Tensor:
    """kaldi.py:436-511 (vtln_warp == 1.0 branch). Returns [num_bins, padded/2]."""
    num_fft_bins = padded / 2
    nyquist = 0.5 * fs
    if high <= 0.0:
        high += nyquist
    fft_bin_width = fs / padded
    mel_low, mel_high = mel_scale_scalar(low), mel_scale_scalar(high)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = mel_scale(fft_bin_width * torch.arange(num_fft_bins)).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down))


def kaldi_fbank(waveform_scaled: Tensor, n_mels=80, fs=16000, frame_length_ms=25.0, frame_shift_ms=10.0) -> Tensor:
    """kaldi.fbank with the arguments WavFrontend passes (wav_frontend.py:171-181): dither=0 (parity
    config), energy_floor=0, hamming, snip_edges=True, remove_dc_offset, preemph 0.97, power, log.
    waveform_scaled: 1-D float32 already multiplied by 32768 (wav_frontend.py:169)."""
    n = waveform_scaled.numel()
    shift = int(fs * frame_shift_ms * 0.001)
    win = int(fs * frame_length_ms * 0.001)
    padded = 1 if win == 0 else 2 ** (win - 1).bit_length()
    if n < win:
        return torch.empty(0, n_mels)
    m = 1 + (n - win) // shift                                     # kaldi.py:_get_strided snip_edges
    x = waveform_scaled.as_strided((m, win), (shift, 1))
    x = x - torch.mean(x, dim=1).unsqueeze(1)                      # remove_dc_offset :183-186
    off = F.pad(x.unsqueeze(0), (1, 0), mode="replicate").squeeze(0)
    x = x - 0.97 * off[:, :-1]                                     # pre-emphasis :193-198
    x = x * torch.hamming_window(win, periodic=False, alpha=0.54, beta=0.46).unsqueeze(0)
    if padded != win:
        x = F.pad(x.unsqueeze(0), (0, padded - win), mode="constant", value=0).squeeze(0)
    spec = torch.fft.rfft(x).abs().pow(2.0)                        # :616-618
    banks = F.pad(get_mel_banks(n_mels, padded, float(fs)), (0, 1), mode="constant", value=0)
    mel = torch.mm(spec, banks.T)                                  # :630
    return torch.max(mel, torch.tensor(EPS_F32)).log()             # :632-633


def apply_lfr(x: Tensor, m: int = 7, n: int = 6) -> Tensor:
    """wav_frontend.py:63-86 restated as the clamped gather it equals (probed):
    out[i] = concat_j x[clamp(n*i - (m-1)//2 + j, 0, T-1)], i < ceil(T/n)."""
    T = x.shape[0]
    T_lfr = (T + n - 1) // n
    idx = (torch.arange(T_lfr).unsqueeze(1) * n - (m - 1) // 2 + torch.arange(m).unsqueeze(0)).clamp_(0, T - 1)
    return x[idx].reshape(T_lfr, m * x.shape[1]).contiguous()


def frontend(wavs: List[Tensor], cmvn: Optional[Tensor], lfr_m=7, lfr_n=6) -> Tuple[Tensor, Tensor]:
    """WavFrontend.forward (wav_frontend.py:149-196). wavs: list of 1-D float32 in [-1,1]."""
    feats, lens = [], []
    for w in wavs:
        n = w.numel()
        mat = kaldi_fbank(w * (1 << 15), frame_length_ms=min(25, n / 16000 * 1000))
        mat = apply_lfr(mat, lfr_m, lfr_n)
        if cmvn is not None:
            mat = (mat + cmvn[0:1]) * cmvn[1:2]
        feats.append(mat)
        lens.append(mat.shape[0])
    pad = torch.nn.utils.rnn.pad_sequence(feats, batch_first=True, padding_value=0.0)
    return pad, torch.tensor(lens, dtype=torch.int32)


# --------------------------------------------------------------------------------------
# Encoder: funasr/models/sanm/encoder.py:392-461, attention.py:216-327
# --------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-12):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)               # transformer/layer_norm.py:13-39


def sinusoid_pe(T: int, depth: int) -> Tensor:
    """transformer/embedding.py:396-432: positions 1..T, concat(sin, cos) halves."""
    pos = torch.arange(1, T + 1)[None, :].type(torch.float32)
    inc = torch.log(torch.tensor([10000], dtype=torch.float32)) / (depth / 2 - 1)
    inv = torch.exp(torch.arange(depth / 2).type(torch.float32) * (-inc))
    st = pos.reshape(1, -1, 1) * inv.reshape(1, 1, -1)
    return torch.cat([torch.sin(st), torch.cos(st)], dim=2)


def fsmn(v: Tensor, w: Tensor, mask_bt1: Tensor) -> Tensor:
    """forward_fsmn attention.py:216-239 / decoder variant :583-631: depthwise conv (k, zero pad (k-1)//2 both sides)."""
    k = w.shape[-1]
    left = (k - 1) // 2
    inp = v * mask_bt1
    x = F.pad(inp.transpose(1, 2), (left, k - 1 - left), value=0.0)
    x = F.conv1d(x, w, None, groups=w.shape[0]).transpose(1, 2)
    return (x + inp) * mask_bt1


def mh_attention(q, k, v, key_mask_b1t, heads: int) -> Tensor:
    """attention.py:288-304 / :760-794: scores, masked_fill(-inf) -> softmax -> masked_fill(0) -> @v, merge heads."""
    B, Tq, D = q.shape
    dk = D // heads
    qh = q.reshape(B, Tq, heads, dk).transpose(1, 2) * dk ** (-0.5)
    kh = k.reshape(B, -1, heads, dk).transpose(1, 2)
    vh = v.reshape(B, -1, heads, dk).transpose(1, 2)
    scores = torch.matmul(qh, kh.transpose(-2, -1))
    m = key_mask_b1t.unsqueeze(1).eq(0)
    attn = torch.softmax(scores.masked_fill(m, -float("inf")), dim=-1).masked_fill(m, 0.0)
    return torch.matmul(attn, vh).transpose(1, 2).contiguous().view(B, Tq, D)


def sanm_attention(u, p: Dict[str, Tensor], pre: str, mask_b1t, heads: int) -> Tensor:
    """MultiHeadedAttentionSANM.forward attention.py:308-327."""
    B, T, _ = u.shape
    qkv = F.linear(u, p[pre + "linear_q_k_v.weight"], p[pre + "linear_q_k_v.bias"])
    D = qkv.shape[-1] // 3
    q, k, v = torch.split(qkv, D, dim=-1)
    mem = fsmn(v, p[pre + "fsmn_block.weight"], mask_b1t.reshape(B, -1, 1).float())
    ctx = mh_attention(q, k, v, mask_b1t, heads)
    return F.linear(ctx, p[pre + "linear_out.weight"], p[pre + "linear_out.bias"]) + mem


def encoder_layer(x, p, pre, mask_b1t, heads, eps, taps=None) -> Tensor:
    """EncoderLayerSANM.forward encoder.py:72-148 (eval, normalize_before, no concat_after)."""
    in_size = p[pre + "norm1.weight"].numel()
    size = p[pre + "norm2.weight"].numel()
    a = sanm_attention(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "self_attn.", mask_b1t, heads)
    x = x + a if in_size == size else a
    h = F.linear(layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps),
                 p[pre + "feed_forward.w_1.weight"], p[pre + "feed_forward.w_1.bias"])
    return x + F.linear(torch.relu(h), p[pre + "feed_forward.w_2.weight"], p[pre + "feed_forward.w_2.bias"])


def encoder(feats: Tensor, lens: Tensor, p: Dict[str, Tensor], enc_layers: int, heads=4, eps=1e-12,
            taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """SANMEncoder.forward encoder.py:392-461 (input_layer='pe')."""
    B, T, Din = feats.shape
    mask = (torch.arange(T)[None, :] < lens[:, None].long())[:, None, :]   # ~make_pad_mask
    D = p["encoder.after_norm.weight"].numel()
    x = feats * D ** 0.5
    x = x + sinusoid_pe(T, Din)
    x = encoder_layer(x, p, "encoder.encoders0.0.", mask, heads, eps)
    if taps is not None:
        taps["enc_l0"] = x
    for i in range(enc_layers - 1):
        x = encoder_layer(x, p, "encoder.encoders.%d." % i, mask, heads, eps)
    x = layer_norm(x, p["encoder.after_norm.weight"], p["encoder.after_norm.bias"], eps)
    return x, mask.squeeze(1).sum(1).to(torch.int32)


# --------------------------------------------------------------------------------------
# CIF predictor: funasr/models/paraformer/cif_predictor.py:253-314, 414-446, 818-908
# --------------------------------------------------------------------------------------
def cif_alphas(enc: Tensor, mask_b1t: Tensor, p, smooth=1.0, noise=0.0) -> Tensor:
    """cif_predictor.py:273-287."""
    q = F.pad(enc.transpose(1, 2), (1, 1), value=0.0)
    out = torch.relu(F.conv1d(q, p["predictor.cif_conv1d.weight"], p["predictor.cif_conv1d.bias"])).transpose(1, 2)
    out = F.linear(out, p["predictor.cif_output.weight"], p["predictor.cif_output.bias"])
    al = torch.relu(torch.sigmoid(out) * smooth - noise)
    return (al * mask_b1t.transpose(-1, -2).float()).squeeze(-1)


def cif_tail(hidden: Tensor, alphas: Tensor, mask_bt: Tensor, tail_threshold: float):
    """tail_process_fn cif_predictor.py:414-446 (mask given)."""
    b, t, d = hidden.shape
    z = torch.zeros((b, 1), dtype=torch.float32)
    m = torch.cat([torch.ones_like(z), mask_bt], dim=1) - torch.cat([mask_bt, z], dim=1)
    alphas = torch.add(torch.cat([alphas, z], dim=1), m * tail_threshold)
    hidden = torch.cat([hidden, torch.zeros((b, 1, d), dtype=hidden.dtype)], dim=1)
    return hidden, alphas, torch.floor(alphas.sum(dim=-1))


def cif_fires(alphas: Tensor):
    """cif_wo_hidden_v1 cif_predictor.py:818-850."""
    ps = torch.cumsum(alphas, dim=1, dtype=torch.float64).to(torch.float32)
    psf = torch.floor(ps)
    dis = torch.floor(torch.roll(ps, 1, dims=1))
    dis[:, 0] = 0
    fire_idxs = (psf - dis) > 0
    fires = torch.zeros_like(alphas)
    fires[fire_idxs] = 1
    return fires + ps - psf, fire_idxs


def cif_v1(hidden: Tensor, alphas: Tensor):
    """cif_v1 cif_predictor.py:853-908, restated per row (identical arithmetic order for every emitted
    frame: ((PH[t_k] - PH[t_{k-1}]) + rem_{k-1} h_{k-1}) - rem_k h_k), with the reference's
    undefined case (a row with zero fires inside a batch that has fires) defined as 'zero tokens'."""
    fires, fire_idxs = cif_fires(alphas)
    B, T, D = hidden.shape
    max_label_len = int(torch.round(alphas.sum(-1)).int().max())
    out = torch.zeros(B, max_label_len, D, dtype=hidden.dtype)
    if fire_idxs.sum() == 0:
        return out, fires
    PH = torch.cumsum(alphas.unsqueeze(-1).repeat((1, 1, D)) * hidden, dim=1)
    rem = fires - torch.floor(fires)
    for b in range(B):
        idx = torch.nonzero(fire_idxs[b]).flatten()
        if idx.numel() == 0:
            continue
        fr = PH[b, idx]
        rf = rem[b, idx].unsqueeze(-1).repeat((1, D)) * hidden[b, idx]
        sfr = torch.roll(fr, 1, dims=0)
        srf = torch.roll(rf, 1, dims=0)
        sfr[0] = 0
        srf[0] = 0
        e = fr - sfr + srf - rf
        n = min(e.shape[0], max_label_len)
        out[b, :n] = e[:n]
    return out, fires


def predictor(enc: Tensor, enc_lens: Tensor, p, tail_threshold=0.45):
    """Paraformer.calc_predictor model.py:315-329 -> CifPredictorV2.forward (inference branch)."""
    B, T, _ = enc.shape
    mask = (torch.arange(T)[None, :] < enc_lens[:, None].long())[:, None, :]
    al = cif_alphas(enc, mask, p)
    hidden, al2, token_num = cif_tail(enc, al, mask.squeeze(1).float(), tail_threshold)
    emb, peaks = cif_v1(hidden, al2)
    n_int = int(torch.max(token_num).type(torch.int32).item())
    return emb[:, :n_int, :], token_num, al2, peaks


# --------------------------------------------------------------------------------------
# Decoder: funasr/models/paraformer/decoder.py:78-121, 397-449
# --------------------------------------------------------------------------------------
def dec_ffn(x, p, pre, eps):
    """PositionwiseFeedForwardDecoderSANM sanm/positionwise_feed_forward.py:12-33."""
    h = torch.relu(F.linear(x, p[pre + "w_1.weight"], p[pre + "w_1.bias"]))
    return F.linear(layer_norm(h, p[pre + "norm.weight"], p[pre + "norm.bias"], eps), p[pre + "w_2.weight"])


def decoder(enc, enc_lens, emb, tok_lens, p, dec_layers: int, heads=4, eps=1e-12, taps=None):
    """ParaformerSANMDecoder.forward decoder.py:397-449 -> logits [B, N, V]."""
    B, N, D = emb.shape
    T = enc.shape[1]
    tgt_mask = (torch.arange(N)[None, :] < tok_lens[:, None].long()).float()[:, :, None]
    mem_mask = (torch.arange(T)[None, :] < enc_lens[:, None].long()).float()[:, None, :]
    x = emb
    for i in range(dec_layers):
        pre = "decoder.decoders.%d." % i
        r = x
        t = dec_ffn(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "feed_forward.", eps)
        t = layer_norm(t, p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
        x = r + fsmn(t, p[pre + "self_attn.fsmn_block.weight"], tgt_mask)
        r = x
        y = layer_norm(x, p[pre + "norm3.weight"], p[pre + "norm3.bias"], eps)
        q = F.linear(y, p[pre + "src_attn.linear_q.weight"], p[pre + "src_attn.linear_q.bias"])
        kv = F.linear(enc, p[pre + "src_attn.linear_k_v.weight"], p[pre + "src_attn.linear_k_v.bias"])
        k, v = torch.split(kv, D, dim=-1)
        ctx = mh_attention(q, k, v, mem_mask, heads)
        x = r + F.linear(ctx, p[pre + "src_attn.linear_out.weight"], p[pre + "src_attn.linear_out.bias"])
        if taps is not None and i == 0:
            taps["dec_l0"] = x
    pre = "decoder.decoders3.0."
    x = dec_ffn(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "feed_forward.", eps)
    h = layer_norm(x, p["decoder.after_norm.weight"], p["decoder.after_norm.bias"], eps)
    return F.linear(h, p["decoder.output_layer.weight"], p["decoder.output_layer.bias"])


def greedy_ids(logp: Tensor, tok_lens: Tensor, sos=1, eos=2, blank=0) -> List[List[int]]:
    """paraformer/model.py:628-666 greedy branch with tokenizer=None -> token_int lists."""
    res = []
    for i in range(logp.shape[0]):
        ys = logp[i, : int(tok_lens[i])].argmax(dim=-1).tolist()
        res.append([t for t in ys if t not in (eos, sos, blank)])
    return res


def paraformer_forward(wavs: List[Tensor], p: Dict[str, Tensor], cmvn: Optional[Tensor], enc_layers: int,
                       dec_layers: int, heads: int = 4, eps: float = 1e-12, tail_threshold: float = 0.45,
                       want_taps: bool = False):
    """Paraformer.inference model.py:534-697 from waveforms to greedy ids (plus stage taps)."""
    taps = {} if want_taps else None
    with torch.no_grad():
        feats, flens = frontend(wavs, cmvn)
        enc, elens = encoder(feats, flens, p, enc_layers, heads, eps, taps)
        emb, token_num, alphas, peaks = predictor(enc, elens, p, tail_threshold)
        tok = token_num.round().long()
        out = {"feats": feats, "feat_lens": flens, "enc": enc, "alphas": alphas, "peaks": peaks, "token_num": tok.to(torch.int32),
               "acoustic": emb}
        if int(tok.max()) < 1:
            out.update(ids=[[] for _ in wavs], logp=None)
            return out
        logits = decoder(enc, elens, emb, tok, p, dec_layers, heads, eps, taps)
        logp = torch.log_softmax(logits, dim=-1)
        out.update(logp=logp, ids=greedy_ids(logp, tok))
        if taps:
            out.update(taps)
    return out


# --------------------------------------------------------------------------------------
# SenseVoiceSmall (BASELINE config 4): funasr/models/sense_voice/model.py:623-656 (encoder), :918-1034 (inference)
# --------------------------------------------------------------------------------------
def sensevoice_encoder(x: Tensor, lens: Tensor, p: Dict[str, Tensor], enc_layers: int, tp_layers: int, heads=4, eps=1e-5):
    """SenseVoiceEncoderSmall.forward model.py:623-656: x*sqrt(512)+PE, encoders0+encoders, after_norm, tp_encoders, tp_norm."""
    B, T, Din = x.shape
    mask = (torch.arange(T)[None, :] < lens[:, None].long())[:, None, :]
    D = p["encoder.after_norm.weight"].numel()
    x = x * D ** 0.5
    x = x + sinusoid_pe(T, Din)
    x = encoder_layer(x, p, "encoder.encoders0.0.", mask, heads, eps)
    for i in range(enc_layers - 1):
        x = encoder_layer(x, p, "encoder.encoders.%d." % i, mask, heads, eps)
    x = layer_norm(x, p["encoder.after_norm.weight"], p["encoder.after_norm.bias"], eps)
    for i in range(tp_layers):
        x = encoder_layer(x, p, "encoder.tp_encoders.%d." % i, mask, heads, eps)
    x = layer_norm(x, p["encoder.tp_norm.weight"], p["encoder.tp_norm.bias"], eps)
    return x, mask.squeeze(1).sum(1).to(torch.int32)


def sensevoice_forward(wavs: List[Tensor], p: Dict[str, Tensor], cmvn: Optional[Tensor], enc_layers: int, tp_layers: int,
                       language_id: int = 0, textnorm_id: int = 15, heads: int = 4, eps: float = 1e-5, blank: int = 0):
    """SenseVoiceSmall.inference model.py:918-1034: frontend -> prepend [language, event(1), emo(2), textnorm] query frames
    (:971-995) -> encoder -> ctc.log_softmax (:1003) -> argmax -> unique_consecutive -> drop blank (:1015-1025)."""
    with torch.no_grad():
        feats, flens = frontend(wavs, cmvn)
        emb = p["embed.weight"]
        q = torch.stack([emb[language_id], emb[1], emb[2], emb[textnorm_id]])[None].repeat(feats.shape[0], 1, 1)
        x = torch.cat([q, feats], dim=1)
        lens = flens + 4
        enc, elens = sensevoice_encoder(x, lens, p, enc_layers, tp_layers, heads, eps)
        logp = torch.log_softmax(F.linear(enc, p["ctc.ctc_lo.weight"], p["ctc.ctc_lo.bias"]), dim=2)
        ids = []
        for i in range(enc.shape[0]):
            y = torch.unique_consecutive(logp[i, : int(elens[i])].argmax(dim=-1), dim=-1)
            ids.append(y[y != blank].tolist())
    return {"feats": feats, "feat_lens": flens, "enc": enc, "enc_lens": elens, "logp": logp, "ids": ids}


# --------------------------------------------------------------------------------------
# ContextualParaformer (BASELINE config 5): funasr/models/contextual_paraformer/model.py:331-385, decoder.py:293-352
# --------------------------------------------------------------------------------------
def lstm_last_hidden(x: Tensor, lengths: List[int], p: Dict[str, Tensor], pre="bias_encoder.") -> Tensor:
    """1-layer batch_first nn.LSTM over packed sequences -> h_n [N, H] (model.py:360-372). Gate order i,f,g,o."""
    wih, whh = p[pre + "weight_ih_l0"], p[pre + "weight_hh_l0"]
    bih, bhh = p[pre + "bias_ih_l0"], p[pre + "bias_hh_l0"]
    N, _, H = x.shape[0], x.shape[1], whh.shape[1]
    out = torch.zeros(N, H)
    for n in range(N):
        h, c = torch.zeros(H), torch.zeros(H)
        for t in range(lengths[n]):
            gates = F.linear(x[n, t], wih, bih) + F.linear(h, whh, bhh)
            i, f, g, o = gates.chunk(4)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
        out[n] = h
    return out


def hotword_embeddings(hw_list: List[List[int]], p: Dict[str, Tensor]) -> Tensor:
    """bias_embed -> LSTM -> last hidden state per hotword: [Nhw, 512] (model.py:350-372)."""
    lens = [len(h) for h in hw_list]
    pad = torch.zeros(len(hw_list), max(lens), dtype=torch.long)
    for i, h in enumerate(hw_list):
        pad[i, : len(h)] = torch.tensor(h)
    return lstm_last_hidden(F.embedding(pad, p["bias_embed.weight"]), lens, p)


def contextual_decoder(enc, enc_lens, emb, tok_lens, hw_embed, p, dec_layers: int, heads=4, eps=1e-12, clas_scale=1.0):
    """ContextualParaformerDecoder.forward decoder.py:293-352 -> logits [B, N, V]."""
    B, N, D = emb.shape
    T = enc.shape[1]
    tgt_mask = (torch.arange(N)[None, :] < tok_lens[:, None].long()).float()[:, :, None]
    mem_mask = (torch.arange(T)[None, :] < enc_lens[:, None].long()).float()[:, None, :]

    def layer(x, pre):
        r = x
        t = dec_ffn(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "feed_forward.", eps)
        t = layer_norm(t, p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
        x_self = r + fsmn(t, p[pre + "self_attn.fsmn_block.weight"], tgt_mask)
        y = layer_norm(x_self, p[pre + "norm3.weight"], p[pre + "norm3.bias"], eps)
        q = F.linear(y, p[pre + "src_attn.linear_q.weight"], p[pre + "src_attn.linear_q.bias"])
        kv = F.linear(enc, p[pre + "src_attn.linear_k_v.weight"], p[pre + "src_attn.linear_k_v.bias"])
        k, v = torch.split(kv, D, dim=-1)
        x_src = F.linear(mh_attention(q, k, v, mem_mask, heads), p[pre + "src_attn.linear_out.weight"], p[pre + "src_attn.linear_out.bias"])
        return x_self + x_src, x_self, x_src

    x = emb
    for i in range(dec_layers - 1):
        x, _, _ = layer(x, "decoder.decoders.%d." % i)
    _, x_self, x_src = layer(x, "decoder.last_decoder.")
    # bias decoder: cross attention over the hotword embeddings (same list for every utterance)
    mem = hw_embed[None].repeat(B, 1, 1)
    pre = "decoder.bias_decoder."
    y = layer_norm(x_self, p[pre + "norm3.weight"], p[pre + "norm3.bias"], eps)
    q = F.linear(y, p[pre + "src_attn.linear_q.weight"], p[pre + "src_attn.linear_q.bias"])
    kv = F.linear(mem, p[pre + "src_attn.linear_k_v.weight"], p[pre + "src_attn.linear_k_v.bias"])
    k, v = torch.split(kv, D, dim=-1)
    cmask = torch.ones(B, 1, mem.shape[1])
    cx = F.linear(mh_attention(q, k, v, cmask, heads), p[pre + "src_attn.linear_out.weight"], p[pre + "src_attn.linear_out.bias"])
    cat = torch.cat([x_src, cx * clas_scale], dim=2)
    x = x_self + F.conv1d(cat.transpose(1, 2), p["decoder.bias_output.weight"]).transpose(1, 2)
    pre = "decoder.decoders3.0."
    x = dec_ffn(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "feed_forward.", eps)
    h = layer_norm(x, p["decoder.after_norm.weight"], p["decoder.after_norm.bias"], eps)
    return F.linear(h, p["decoder.output_layer.weight"], p["decoder.output_layer.bias"])


def contextual_forward(wavs, p, cmvn, enc_layers: int, dec_layers: int, hw_list, heads=4, eps=1e-12, tail_threshold=0.45):
    """ContextualParaformer.inference greedy path (model.py:387-520) with a hotword list (token ids)."""
    with torch.no_grad():
        feats, flens = frontend(wavs, cmvn)
        enc, elens = encoder(feats, flens, p, enc_layers, heads, eps)
        emb, token_num, alphas, peaks = predictor(enc, elens, p, tail_threshold)
        tok = token_num.round().long()
        hw = hotword_embeddings(hw_list, p)
        logits = contextual_decoder(enc, elens, emb, tok, hw, p, dec_layers, heads, eps)
        logp = torch.log_softmax(logits, dim=-1)
    return {"enc": enc, "token_num": tok.to(torch.int32), "acoustic": emb, "hw_embed": hw, "logp": logp, "ids": greedy_ids(logp, tok)}


# --------------------------------------------------------------------------------------
# BiCifParaformer: funasr/models/bicif_paraformer/{cif_predictor.py,model.py}
# --------------------------------------------------------------------------------------
def cif_loop(hidden: Tensor, alphas: Tensor, threshold: float = 1.0):
    """`cif` bicif_paraformer/cif_predictor.py:37-84: sequential fp32 integrate-and-fire; returns (frames padded to the
    largest fire count, fires)."""
    B, T, D = hidden.shape
    integrate = torch.zeros(B)
    frame = torch.zeros(B, D)
    fires, frames = [], []
    for t in range(T):
        alpha = alphas[:, t]
        completion = torch.ones(B) - integrate
        integrate = integrate + alpha
        fires.append(integrate)
        fire = integrate >= threshold
        integrate = torch.where(fire, integrate - torch.ones(B), integrate)
        cur = torch.where(fire, completion, alpha)
        rem = alpha - cur
        frame = frame + cur[:, None] * hidden[:, t, :]
        frames.append(frame)
        frame = torch.where(fire[:, None], rem[:, None] * hidden[:, t, :], frame)
    fires = torch.stack(fires, 1)
    frames = torch.stack(frames, 1)
    rows = [frames[b][fires[b] >= threshold] for b in range(B)]
    n = max(int(r.shape[0]) for r in rows)
    out = torch.zeros(B, n, D)
    for b, r in enumerate(rows):
        out[b, : r.shape[0]] = r
    return out, fires


def predictor_v3(enc: Tensor, enc_lens: Tensor, p, tail_threshold=0.45, threshold=1.0):
    """BiCifParaformer.calc_predictor model.py:162-175 -> CifPredictorV3.forward (inference branch, :219-298)."""
    B, T, _ = enc.shape
    mask = (torch.arange(T)[None, :] < enc_lens[:, None].long())[:, None, :]
    al = cif_alphas(enc, mask, p)
    hidden, al2, token_num = cif_tail(enc, al, mask.squeeze(1).float(), tail_threshold)
    emb, fires = cif_loop(hidden, al2, threshold)
    n_int = int(torch.max(token_num).type(torch.int32).item())
    if emb.shape[1] < n_int:
        emb = F.pad(emb, (0, 0, 0, n_int - emb.shape[1]))
    return emb[:, :n_int, :], token_num, al2, fires


def cif_wo_hidden_loop(alphas: Tensor, threshold: float) -> Tensor:
    """`cif_wo_hidden` bicif_paraformer/cif_predictor.py:87-117."""
    B, T = alphas.shape
    integrate = torch.zeros(B)
    fires = []
    for t in range(T):
        integrate = integrate + alphas[:, t]
        fires.append(integrate)
        integrate = torch.where(integrate >= threshold, integrate - torch.ones(B) * threshold, integrate)
    return torch.stack(fires, 1)


def upsample_timestamp(enc: Tensor, enc_lens: Tensor, token_num: Tensor, p, smooth2=0.25, noise2=0.01, threshold=1.0, times=3):
    """CifPredictorV3.get_upsample_timestamp :300-352 with upsample_type "cnn_blstm", use_cif1_cnn False
    -> (us_alphas [B, 3T], us_peaks [B, 3T])."""
    B, T, D = enc.shape
    up = F.conv_transpose1d(enc.transpose(1, 2), p["predictor.upsample_cnn.weight"], p["predictor.upsample_cnn.bias"], stride=times)
    lstm = torch.nn.LSTM(D, D, 1, bias=True, batch_first=True, dropout=0.0, bidirectional=True)
    lstm.load_state_dict({k[len("predictor.blstm."):]: v for k, v in p.items() if k.startswith("predictor.blstm.")})
    with torch.no_grad():
        o2, _ = lstm(up.transpose(1, 2))
    a2 = torch.sigmoid(F.linear(o2, p["predictor.cif_output2.weight"], p["predictor.cif_output2.bias"]))
    a2 = torch.relu(a2 * smooth2 - noise2)
    mask = (torch.arange(T)[None, :] < enc_lens[:, None].long())[:, None, :]
    mask2 = mask.repeat(1, times, 1).transpose(-1, -2).reshape(B, -1).unsqueeze(-1)
    a2 = (a2 * mask2).squeeze(-1)
    tot = a2.sum(-1)
    a2 = a2 * (token_num / tot)[:, None].repeat(1, a2.size(1))
    return a2, cif_wo_hidden_loop(a2, threshold - 1e-4)


def bicif_forward(wavs: List[Tensor], p: Dict[str, Tensor], cmvn: Optional[Tensor], enc_layers: int, dec_layers: int, heads: int = 4,
                  eps: float = 1e-12, tail_threshold: float = 0.45):
    """BiCifParaformer.inference model.py:271-428 (greedy, tokenizer=None) plus the raw CIF timestamps it would attach."""
    with torch.no_grad():
        feats, flens = frontend(wavs, cmvn)
        enc, elens = encoder(feats, flens, p, enc_layers, heads, eps, None)
        emb, token_num, alphas, fires = predictor_v3(enc, elens, p, tail_threshold)
        tok = token_num.round().long()
        out = {"enc": enc, "enc_lens": elens, "alphas": alphas, "peaks": fires, "token_num": tok.to(torch.int32), "acoustic": emb}
        if int(tok.max()) < 1:
            out.update(ids=[[] for _ in wavs])
            return out
        logits = decoder(enc, elens, emb, tok, p, dec_layers, heads, eps, None)
        logp = torch.log_softmax(logits, dim=-1)
        us_alphas, us_peaks = upsample_timestamp(enc, elens, tok, p)
        out.update(logp=logp, ids=greedy_ids(logp, tok), us_alphas=us_alphas, us_peaks=us_peaks)
    return out



# --------------------------------------------------------------------------------------
# torch's CPU fp32 row sum, restated step by step (ATen/native/cpu/SumKernel.cpp: cascade_sum -> vectorized_inner_sum ->
# row_sum -> multi_row_sum).  The CIF token count is floor(alphas.sum(-1)) (cif_predictor.py:443-444), so the SUMMATION ORDER
# decides an integer outcome; the CUDA kernel (csrc/cif.cu: torch_row_sum_f32) follows the same order and
# tests/test_oracle_golden.py::test_torch_row_sum_emulation pins this restatement to torch.sum itself.
# --------------------------------------------------------------------------------------
def _multi_row_sum(R):
    """R [size, 4, L] fp32 -> 4 accumulators [L]: 4-level cascade, level 0 flushed every 16 rows (level_power 4)."""
    import numpy as np
    size, L = R.shape[0], R.shape[2]
    acc = np.zeros((4, 4, L), np.float32)
    i = 0
    while i + 16 <= size:
        for _ in range(16):
            acc[0] = acc[0] + R[i]
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = 0
            if (i & (15 << (4 * j))) != 0:
                break
    while i < size:
        acc[0] = acc[0] + R[i]
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    return [acc[0][k].copy() for k in range(4)]


def torch_row_sum_f32(x, lanes: int = 8):
    """x: 1-D fp32 array -> np.float32 equal to torch.from_numpy(x[None]).sum(-1) on x86 CPUs (8 lanes under every capability)."""
    import numpy as np
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    L = lanes if n >= lanes else 1                       # rows shorter than a vector take scalar_inner_sum: same scheme, one lane
    vec = n // L
    V = x[: vec * L].reshape(vec, L)
    ilp = vec // 4
    ps = _multi_row_sum(V[: ilp * 4].reshape(ilp, 4, L))
    for i in range(ilp * 4, vec):
        ps[0] = ps[0] + V[i]
    for k in range(1, 4):
        ps[0] = ps[0] + ps[k]
    f = np.float32(0)
    for k in range(vec * L, n):
        f = np.float32(f + x[k])
    for k in range(L):
        f = np.float32(f + ps[0][k])
    return f

# --------------------------------------------------------------------------------------
# SeacoParaformer: funasr/models/seaco_paraformer/model.py (greedy inference with hotwords, ASF)
# --------------------------------------------------------------------------------------
def sanm_decoder_layers(x, x_lens, mem, mem_lens, p, prefix: str, layers, heads=4, eps=1e-12, attn_of: Optional[int] = None):
    """DecoderLayerSANM.forward (paraformer/decoder.py:78-121) for `layers` of the decoder stored under `prefix`; with attn_of = i
    the cross-attention matrix of layer i is returned instead (get_attn_mat :123-146) after running the layers before it."""
    B, N, D = x.shape
    T = mem.shape[1]
    x_mask = (torch.arange(N)[None, :] < x_lens[:, None].long()).float()[:, :, None]
    mem_mask = (torch.arange(T)[None, :] < mem_lens[:, None].long()).float()[:, None, :]
    for i in layers:
        pre = "%sdecoders.%d." % (prefix, i)
        r = x
        t = dec_ffn(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "feed_forward.", eps)
        t = layer_norm(t, p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
        x = r + fsmn(t, p[pre + "self_attn.fsmn_block.weight"], x_mask)
        r = x
        y = layer_norm(x, p[pre + "norm3.weight"], p[pre + "norm3.bias"], eps)
        q = F.linear(y, p[pre + "src_attn.linear_q.weight"], p[pre + "src_attn.linear_q.bias"])
        kv = F.linear(mem, p[pre + "src_attn.linear_k_v.weight"], p[pre + "src_attn.linear_k_v.bias"])
        k, v = torch.split(kv, D, dim=-1)
        if attn_of is not None and i == attn_of:
            dk = D // heads
            qh = q.view(B, N, heads, dk).transpose(1, 2) * dk ** (-0.5)
            kh = k.view(B, T, heads, dk).transpose(1, 2)
            sc = torch.matmul(qh, kh.transpose(-2, -1)).masked_fill(mem_mask[:, None].eq(0), -float("inf"))
            return torch.softmax(sc, dim=-1).masked_fill(mem_mask[:, None].eq(0), 0.0)      # sanm/attention.py:781-793
        x = r + F.linear(mh_attention(q, k, v, mem_mask, heads), p[pre + "src_attn.linear_out.weight"], p[pre + "src_attn.linear_out.bias"])
    return x


def sanm_decoder_hidden(x, x_lens, mem, mem_lens, p, prefix: str, n_layers: int, heads=4, eps=1e-12):
    """ParaformerSANMDecoder.forward with return_hidden (decoder.py:397-449): decoders -> decoders3 -> after_norm."""
    x = sanm_decoder_layers(x, x_lens, mem, mem_lens, p, prefix, range(n_layers), heads, eps)
    pre = prefix + "decoders3.0."
    x = dec_ffn(layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps), p, pre + "feed_forward.", eps)
    return layer_norm(x, p[prefix + "after_norm.weight"], p[prefix + "after_norm.bias"], eps)


def seaco_hotword_representation(hw_list: List[List[int]], p: Dict[str, Tensor]) -> Tensor:
    """_hotword_representation model.py:384-397: decoder.embed -> 2-layer LSTM over the PADDED batch (no packing) -> the output at
    each hotword's last token."""
    lens = [len(h) for h in hw_list]
    pad = torch.zeros(len(hw_list), max(lens), dtype=torch.long)
    for i, h in enumerate(hw_list):
        pad[i, : len(h)] = torch.tensor(h)
    lstm = torch.nn.LSTM(512, 512, 2, batch_first=True)
    lstm.load_state_dict({k[len("bias_encoder."):]: v for k, v in p.items() if k.startswith("bias_encoder.")})
    with torch.no_grad():
        out, _ = lstm(F.embedding(pad, p["decoder.embed.0.weight"]))
    return out[torch.arange(len(hw_list)), torch.tensor(lens) - 1]


def seaco_decode_with_asf(enc, enc_lens, emb, tok, hw_list, p, dec_layers: int, no_bias: int, nfilter=50, seaco_weight=1.0, heads=4,
                          eps=1e-12, seaco_layers=6):
    """_seaco_decode_with_ASF model.py:271-382 -> merged log-probabilities [B, N, V] (plus the taps the tests compare)."""
    hidden = sanm_decoder_hidden(emb, tok, enc, enc_lens, p, "decoder.", dec_layers, heads, eps)
    dec_pred = torch.log_softmax(F.linear(hidden, p["decoder.output_layer.weight"], p["decoder.output_layer.bias"]), dim=-1)
    B = enc.shape[0]
    selected = seaco_hotword_representation(hw_list, p)
    selected_all = selected
    ctx = selected[None].repeat(B, 1, 1)
    n_hw = ctx.shape[1]
    picked = None
    if 0 < nfilter < n_hw:                                             # ASF: keep the nfilter hotwords the FIRST utterance attends to most
        attn = sanm_decoder_layers(hidden, tok, ctx, torch.full((B,), n_hw), p, "seaco_decoder.", range(seaco_layers), heads, eps,
                                   attn_of=seaco_layers - 1)
        scores = attn[0].sum(0).sum(0)
        picked = torch.topk(scores, min(nfilter, n_hw - 1))[1].tolist() + [len(hw_list) - 1]
        selected = selected[picked]
        ctx = selected[None].repeat(B, 1, 1)
        n_hw = ctx.shape[1]
    hl = torch.full((B,), n_hw)
    cif_att = sanm_decoder_hidden(emb, tok, ctx, hl, p, "seaco_decoder.", seaco_layers, heads, eps)
    dec_att = sanm_decoder_hidden(hidden, tok, ctx, hl, p, "seaco_decoder.", seaco_layers, heads, eps)
    dha_pred = torch.log_softmax(F.linear(cif_att + dec_att, p["hotword_output_layer.weight"], p["hotword_output_layer.bias"]), dim=-1)
    lmbd = torch.full((B,), float(seaco_weight))
    mask = (dha_pred.max(-1)[1] == no_bias).int().unsqueeze(-1)
    a, b = ((1 - lmbd) / lmbd).reshape(-1, 1, 1), (1 / lmbd).reshape(-1, 1, 1)
    mask = (mask + a) / b
    return dec_pred * mask + dha_pred * (1 - mask), {"dec_hidden": hidden, "hw_selected": selected, "hw_selected_all": selected_all,
                                                            "asf_picked": picked, "dha_pred": dha_pred}


def seaco_forward(wavs: List[Tensor], p: Dict[str, Tensor], cmvn: Optional[Tensor], enc_layers: int, dec_layers: int, hw_list, no_bias: int,
                  nfilter: int = 50, heads: int = 4, eps: float = 1e-12, tail_threshold: float = 0.45):
    """SeacoParaformer.inference model.py:422-581 (greedy, tokenizer=None) with a hotword id list."""
    with torch.no_grad():
        feats, flens = frontend(wavs, cmvn)
        enc, elens = encoder(feats, flens, p, enc_layers, heads, eps, None)
        emb, token_num, alphas, fires = predictor_v3(enc, elens, p, tail_threshold)
        tok = token_num.round().long()
        merged, taps = seaco_decode_with_asf(enc, elens, emb, tok, hw_list, p, dec_layers, no_bias, nfilter, 1.0, heads, eps)
        us_alphas, us_peaks = upsample_timestamp(enc, elens, tok, p)
    out = {"enc_lens": elens, "token_num": tok.to(torch.int32), "merged": merged, "ids": greedy_ids(merged, tok), "us_alphas": us_alphas,
           "us_peaks": us_peaks}
    out.update(taps)
    return out
