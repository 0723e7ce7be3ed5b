"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, funasr 1.4.3) on CPU.

Run here (the build container), never on the GPU box:   python oracle/make_golden.py
The reference is driven through its own plugin surface: AutoModel(model="Paraformer", model_conf=...,
init_param=<synthetic model.pt>) -> model.inference(..., tokenizer=None) for the greedy ids
(paraformer/model.py:694) and model.encode / calc_predictor / cal_decoder_with_predictor for stage taps
(SURVEY.md App. B).  Weights/waveforms come from funasr_b200/synth.py (seeded), so only outputs are stored.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from funasr_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (cfg, weight seed, [(n_samples, wav seed, kind)], use_cmvn)
    "tiny_ragged3": (synth.PARAFORMER_TINY, 3, [(48000, 1, "speechlike"), (27200, 2, "noise"), (38437, 3, "speechlike")], True),
    "tiny_single": (synth.PARAFORMER_TINY, 3, [(16000, 4, "speechlike")], False),
    "large_ragged2": (synth.PARAFORMER_LARGE, 0, [(480000, 0, "speechlike"), (196800, 5, "speechlike")], True),
}


def write_cmvn_file(path, cmvn):
    """Kaldi-nnet text layout parsed by load_cmvn (wav_frontend.py:15-43)."""
    d = cmvn.shape[1]
    with open(path, "w") as f:
        f.write("<Nnet>\n<Splice> %d %d\n[ 0 ]\n<AddShift> %d %d\n" % (d, d, d, d))
        f.write("<LearnRateCoef> 0 [ " + " ".join("%.9g" % v for v in cmvn[0].tolist()) + " ]\n")
        f.write("<Rescale> %d %d\n" % (d, d))
        f.write("<LearnRateCoef> 0 [ " + " ".join("%.9g" % v for v in cmvn[1].tolist()) + " ]\n</Nnet>\n")


def build_reference_model(cfg, seed, cmvn_file, tmp):
    from funasr import AutoModel
    sd = synth.make_state_dict(cfg, seed)
    pt = os.path.join(tmp, "model_%d_%d.pt" % (cfg.enc_layers, seed))
    torch.save(sd, pt)
    tokens = ["<blank>", "<s>", "</s>"] + ["t%d" % i for i in range(cfg.vocab - 4)] + ["<unk>"]
    am = AutoModel(
        model="Paraformer",
        model_conf=dict(ctc_weight=0.0, lsm_weight=0.1, length_normalized_loss=True, predictor_weight=1.0,
                        predictor_bias=1, sampling_ratio=0.75),
        encoder="SANMEncoder",
        encoder_conf=dict(output_size=cfg.d_model, attention_heads=cfg.heads, linear_units=cfg.ffn,
                          num_blocks=cfg.enc_layers, dropout_rate=0.1, positional_dropout_rate=0.1,
                          attention_dropout_rate=0.1, input_layer="pe", pos_enc_class="SinusoidalPositionEncoder",
                          normalize_before=True, kernel_size=cfg.kernel, sanm_shfit=0, selfattention_layer_type="sanm"),
        decoder="ParaformerSANMDecoder",
        decoder_conf=dict(attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.dec_layers, dropout_rate=0.1,
                          positional_dropout_rate=0.1, self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1,
                          att_layer_num=cfg.dec_layers, kernel_size=cfg.kernel, sanm_shfit=0),
        predictor="CifPredictorV2",
        predictor_conf=dict(idim=cfg.d_model, threshold=1.0, l_order=1, r_order=1, tail_threshold=cfg.tail_threshold),
        frontend="WavFrontend",
        frontend_conf=dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6,
                           dither=0.0, cmvn_file=cmvn_file),
        tokenizer="CharTokenizer",
        tokenizer_conf=dict(token_list=tokens, unk_symbol="<unk>", split_with_space=True),
        device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt,
    )
    return am


def run_case(name, cfg, wseed, wav_specs, use_cmvn, tmp):
    cmvn_file = None
    if use_cmvn:
        cmvn_file = os.path.join(tmp, "am_%s.mvn" % name)
        write_cmvn_file(cmvn_file, synth.make_cmvn(cfg, seed=1))
    am = build_reference_model(cfg, wseed, cmvn_file, tmp)
    model, frontend = am.model, am.kwargs["frontend"]
    wavs = [synth.make_wav(n, s, k) for (n, s, k) in wav_specs]
    with torch.no_grad():
        res, meta = model.inference(data_in=[w.numpy() for w in wavs], key=["u%d" % i for i in range(len(wavs))],
                                    tokenizer=None, frontend=frontend, device="cpu")
        from funasr.utils.load_utils import extract_fbank
        feats, flens = extract_fbank([w for w in wavs], frontend=frontend)
        enc, elens = model.encode(feats, flens)
        emb, tok, alphas, peaks = model.calc_predictor(enc, elens)
        tokl = tok.round().long()
        logp, _ = model.cal_decoder_with_predictor(enc, elens, emb, tokl)
    ids = [r["token_int"] for r in res]
    top2 = torch.topk(logp, 2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    N = logp.shape[1]
    keep = sorted(set(list(range(min(4, N))) + [N // 2, N - 1]))
    out = dict(
        feats=feats.numpy(), feat_lens=flens.numpy().astype(np.int32),
        enc=enc.numpy().astype(np.float32), alphas=alphas.numpy(), token_num=tokl.numpy().astype(np.int32),
        acoustic=emb.numpy(), logp_rows=np.array(keep, dtype=np.int32), logp_sel=logp[:, keep, :].numpy(),
        logp_max=logp.max(-1).values.numpy(), argmax=logp.argmax(-1).numpy().astype(np.int32), margin=margin.numpy(),
        ids_flat=np.array([t for r in ids for t in r], dtype=np.int32), ids_len=np.array([len(r) for r in ids], dtype=np.int32),
        batch_data_time=np.float64(meta["batch_data_time"]),
    )
    if cfg.enc_layers > 10:   # keep the large fixture small: drop the big dense taps, keep strided samples
        out["feats"] = out["feats"][:, ::7, :]
        out["enc"] = out["enc"][:, ::7, :]
        out["acoustic"] = out["acoustic"][:, ::5, :]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    valid = torch.arange(N)[None, :] < tokl[:, None]
    print("%s: B=%d T=%d tokens=%s min_margin(valid)=%.3e  ids[0][:8]=%s" % (
        name, len(wavs), feats.shape[1], tokl.tolist(), float(margin[valid].min()), ids[0][:8]))


def main():
    ref_shim.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:] or list(CASES)
    with tempfile.TemporaryDirectory() as tmp:
        for name in only:
            cfg, wseed, specs, use_cmvn = CASES[name]
            run_case(name, cfg, wseed, specs, use_cmvn, tmp)
    # a cmvn file fixture in the Kaldi-nnet text layout for the parser test
    write_cmvn_file(os.path.join(GOLD, "am_synth.mvn"), synth.make_cmvn(synth.PARAFORMER_LARGE, seed=1))


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("sensevoice", "contextual", "bicif", "seaco")):
    main()


# ------------------------------------------------------------------------------------------------ SenseVoiceSmall
SV_CASES = {
    "sv_tiny_ragged3": (synth.SENSEVOICE_TINY, 4, [(48000, 11, "speechlike"), (27200, 12, "noise"), (38437, 13, "speechlike")], True),
    "sv_large_single": (synth.SENSEVOICE_SMALL, 1, [(160000, 14, "speechlike")], True),
}


def run_sv_case(name, cfg, wseed, wav_specs, use_cmvn, tmp):
    from funasr import AutoModel
    cmvn_file = os.path.join(tmp, "am_%s.mvn" % name)
    write_cmvn_file(cmvn_file, synth.make_cmvn(synth.PARAFORMER_LARGE, seed=1))
    pt = os.path.join(tmp, "sv_%s.pt" % name)
    torch.save(synth.make_sensevoice_state_dict(cfg, wseed), pt)
    tokens = ["<blank>"] + ["t%d" % i for i in range(cfg.vocab - 2)] + ["<unk>"]
    am = AutoModel(
        model="SenseVoiceSmall", model_conf=dict(length_normalized_loss=True, sos=1, eos=2, ignore_id=-1),
        encoder="SenseVoiceEncoderSmall",
        encoder_conf=dict(output_size=cfg.d_model, attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.enc_layers,
                          tp_blocks=cfg.tp_layers, dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1,
                          input_layer="pe", pos_enc_class="SinusoidalPositionEncoder", normalize_before=True, kernel_size=cfg.kernel,
                          sanm_shfit=0, selfattention_layer_type="sanm"),
        frontend="WavFrontend",
        frontend_conf=dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6, dither=0.0,
                           cmvn_file=cmvn_file),
        tokenizer="CharTokenizer", tokenizer_conf=dict(token_list=tokens, unk_symbol="<unk>", split_with_space=True),
        device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt,
    )
    model, frontend = am.model, am.kwargs["frontend"]

    class Tok:      # SenseVoiceSmall.inference needs tokenizer.decode(token_int) (model.py:1027); record the ids verbatim
        def decode(self, ids):
            return " ".join(str(int(i)) for i in ids)

    wavs = [synth.make_wav(n, s, k) for (n, s, k) in wav_specs]
    with torch.no_grad():
        res, meta = model.inference(data_in=[w.numpy() for w in wavs], key=["u%d" % i for i in range(len(wavs))], tokenizer=Tok(),
                                    frontend=frontend, device="cpu", language="auto", use_itn=False)
        from funasr.utils.load_utils import extract_fbank
        feats, flens = extract_fbank([w for w in wavs], frontend=frontend)
        emb = model.embed.weight
        q = torch.stack([emb[0], emb[1], emb[2], emb[15]])[None].repeat(feats.shape[0], 1, 1)
        enc, elens = model.encoder(torch.cat([q, feats], dim=1), flens + 4)
        logp = model.ctc.log_softmax(enc)
    ids = [[int(t) for t in r["text"].split()] for r in res]
    step = 7 if cfg.enc_layers > 10 else 1
    top2 = torch.topk(logp, 2, dim=-1).values
    rows = sorted(set([0, 1, 4, enc.shape[1] // 2, enc.shape[1] - 1]))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), enc=enc[:, ::step].numpy(), enc_lens=elens.numpy().astype(np.int32),
                        logp_rows=np.array(rows, dtype=np.int32), logp_sel=logp[:, rows, :].numpy(),
                        argmax=logp.argmax(-1).numpy().astype(np.int32), margin=(top2[..., 0] - top2[..., 1]).numpy(),
                        ids_flat=np.array([t for r in ids for t in r], dtype=np.int32), ids_len=np.array([len(r) for r in ids], dtype=np.int32))
    print("%s: B=%d T=%d ctc tokens=%s min margin %.3e ids[0][:8]=%s" % (name, len(wavs), enc.shape[1], [len(r) for r in ids],
                                                                       float((top2[..., 0] - top2[..., 1]).min()), ids[0][:8]))


def main_sv():
    ref_shim.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        for name, (cfg, wseed, specs, use_cmvn) in SV_CASES.items():
            run_sv_case(name, cfg, wseed, specs, use_cmvn, tmp)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sensevoice":
    main_sv()


# ------------------------------------------------------------------------------------------------ ContextualParaformer
CTX_CASES = {
    "ctx_tiny_ragged3": (synth.PARAFORMER_TINY, 6, [(48000, 21, "speechlike"), (27200, 22, "noise"), (38437, 23, "speechlike")], 5),
    "ctx_large_single": (synth.PARAFORMER_LARGE, 2, [(240000, 24, "speechlike")], 32),
}


def run_ctx_case(name, cfg, wseed, wav_specs, n_hot, tmp):
    from funasr import AutoModel
    cmvn_file = os.path.join(tmp, "am_%s.mvn" % name)
    write_cmvn_file(cmvn_file, synth.make_cmvn(cfg, seed=1))
    pt = os.path.join(tmp, "ctx_%s.pt" % name)
    torch.save(synth.make_contextual_state_dict(cfg, wseed), pt)
    tokens = ["<blank>", "<s>", "</s>"] + ["t%d" % i for i in range(cfg.vocab - 4)] + ["<unk>"]
    am = AutoModel(
        model="ContextualParaformer",
        model_conf=dict(ctc_weight=0.0, lsm_weight=0.1, length_normalized_loss=True, predictor_weight=1.0, predictor_bias=1,
                        sampling_ratio=0.75, inner_dim=512),
        encoder="SANMEncoder",
        encoder_conf=dict(output_size=cfg.d_model, attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.enc_layers,
                          dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1, input_layer="pe",
                          pos_enc_class="SinusoidalPositionEncoder", normalize_before=True, kernel_size=cfg.kernel, sanm_shfit=0,
                          selfattention_layer_type="sanm"),
        decoder="ContextualParaformerDecoder",
        decoder_conf=dict(attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.dec_layers, dropout_rate=0.1,
                          positional_dropout_rate=0.1, self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1,
                          att_layer_num=cfg.dec_layers, kernel_size=cfg.kernel, sanm_shfit=0),
        predictor="CifPredictorV2",
        predictor_conf=dict(idim=cfg.d_model, threshold=1.0, l_order=1, r_order=1, tail_threshold=cfg.tail_threshold),
        frontend="WavFrontend",
        frontend_conf=dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6, dither=0.0,
                           cmvn_file=cmvn_file),
        tokenizer="CharTokenizer", tokenizer_conf=dict(token_list=tokens, unk_symbol="<unk>", split_with_space=True),
        device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt,
    )
    model, frontend = am.model, am.kwargs["frontend"]
    hw_list = synth.make_hotwords(n_hot, cfg.vocab, seed=7)
    wavs = [synth.make_wav(n, s, k) for (n, s, k) in wav_specs]
    with torch.no_grad():
        from funasr.utils.load_utils import extract_fbank
        feats, flens = extract_fbank([w for w in wavs], frontend=frontend)
        enc, elens = model.encode(feats, flens)
        emb, tok, alphas, peaks = model.calc_predictor(enc, elens)
        tokl = tok.round().long()
        logp, _ = model.cal_decoder_with_predictor(enc, elens, emb, tokl, hw_list=hw_list, clas_scale=1.0)
        # hotword embeddings exactly as cal_decoder_with_predictor builds them (model.py:360-372)
        from funasr.models.transformer.utils.nets_utils import pad_list
        pad = pad_list([torch.Tensor(i).long() for i in hw_list], 0)
        packed = torch.nn.utils.rnn.pack_padded_sequence(model.bias_embed(pad), [len(i) for i in hw_list], batch_first=True, enforce_sorted=False)
        _, (h_n, _) = model.bias_encoder(packed)
    ids = []
    for i in range(len(wavs)):
        ys = logp[i, : int(tokl[i])].argmax(-1).tolist()
        ids.append([t for t in ys if t not in (0, 1, 2)])
    N = logp.shape[1]
    keep = sorted(set(list(range(min(4, N))) + [N // 2, N - 1]))
    top2 = torch.topk(logp, 2, dim=-1).values
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), token_num=tokl.numpy().astype(np.int32), hw_embed=h_n[0].numpy(),
                        logp_rows=np.array(keep, dtype=np.int32), logp_sel=logp[:, keep, :].numpy(),
                        argmax=logp.argmax(-1).numpy().astype(np.int32), margin=(top2[..., 0] - top2[..., 1]).numpy(),
                        ids_flat=np.array([t for r in ids for t in r], dtype=np.int32), ids_len=np.array([len(r) for r in ids], dtype=np.int32))
    valid = torch.arange(N)[None, :] < tokl[:, None]
    print("%s: B=%d tokens=%s n_hot=%d min margin %.3e ids[0][:6]=%s" % (name, len(wavs), tokl.tolist(), len(hw_list),
                                                                      float((top2[..., 0] - top2[..., 1])[valid].min()), ids[0][:6]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "contextual":
    ref_shim.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        for name, (cfg, wseed, specs, n_hot) in CTX_CASES.items():
            run_ctx_case(name, cfg, wseed, specs, n_hot, tmp)


# ------------------------------------------------------------------------------------------------ BiCifParaformer
BICIF_CASES = {
    "bicif_tiny_ragged3": (synth.PARAFORMER_TINY, 8, [(48000, 31, "speechlike"), (27200, 32, "noise"), (38437, 33, "speechlike")]),
    "bicif_large_single": (synth.PARAFORMER_LARGE, 3, [(160000, 34, "speechlike")]),
}


def run_bicif_case(name, cfg, wseed, wav_specs, tmp):
    import copy
    from funasr import AutoModel
    from funasr.utils.timestamp_tools import ts_prediction_lfr6_standard
    cmvn_file = os.path.join(tmp, "am_%s.mvn" % name)
    write_cmvn_file(cmvn_file, synth.make_cmvn(cfg, seed=1))
    pt = os.path.join(tmp, "bicif_%s.pt" % name)
    torch.save(synth.make_bicif_state_dict(cfg, wseed), pt)
    tokens = ["<blank>", "<s>", "</s>"] + ["t%d" % i for i in range(cfg.vocab - 4)] + ["<unk>"]
    am = AutoModel(
        model="BiCifParaformer",
        model_conf=dict(ctc_weight=0.0, lsm_weight=0.1, length_normalized_loss=True, predictor_weight=1.0, predictor_bias=1, sampling_ratio=0.75),
        encoder="SANMEncoder",
        encoder_conf=dict(output_size=cfg.d_model, attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.enc_layers,
                          dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1, input_layer="pe",
                          pos_enc_class="SinusoidalPositionEncoder", normalize_before=True, kernel_size=cfg.kernel, sanm_shfit=0,
                          selfattention_layer_type="sanm"),
        decoder="ParaformerSANMDecoder",
        decoder_conf=dict(attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.dec_layers, dropout_rate=0.1,
                          positional_dropout_rate=0.1, self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1,
                          att_layer_num=cfg.dec_layers, kernel_size=cfg.kernel, sanm_shfit=0),
        predictor="CifPredictorV3",
        predictor_conf=dict(idim=cfg.d_model, threshold=1.0, l_order=1, r_order=1, tail_threshold=cfg.tail_threshold, smooth_factor2=0.25,
                            noise_threshold2=0.01, upsample_times=3, use_cif1_cnn=False, upsample_type="cnn_blstm"),
        frontend="WavFrontend",
        frontend_conf=dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6, dither=0.0,
                           cmvn_file=cmvn_file),
        tokenizer="CharTokenizer", tokenizer_conf=dict(token_list=tokens, unk_symbol="<unk>", split_with_space=True),
        device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt,
    )
    model, frontend = am.model, am.kwargs["frontend"]
    wavs = [synth.make_wav(n, s, k) for (n, s, k) in wav_specs]
    with torch.no_grad():
        from funasr.utils.load_utils import extract_fbank
        feats, flens = extract_fbank([w for w in wavs], frontend=frontend)
        enc, elens = model.encode(feats, flens)
        emb, tok, alphas, peaks = model.calc_predictor(enc, elens)
        tokl = tok.round().long()
        logp, _ = model.cal_decoder_with_predictor(enc, elens, emb, tokl)
        _, _, us_alphas, us_peaks = model.calc_predictor_timestamp(enc, elens, tokl)
    ids, stamps = [], []
    for i in range(len(wavs)):
        ys = logp[i, : int(tokl[i])].argmax(-1).tolist()
        ids.append([t for t in ys if t not in (0, 1, 2)])
        n = int(elens[i]) * 3
        _, st = ts_prediction_lfr6_standard(us_alphas[i][:n].clone(), us_peaks[i][:n].clone(), copy.copy([tokens[t] for t in ids[-1]]), vad_offset=0)
        stamps.append(st)
    N = logp.shape[1]
    keep = sorted(set(list(range(min(4, N))) + [N // 2, N - 1]))
    top2 = torch.topk(logp, 2, dim=-1).values
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), token_num=tokl.numpy().astype(np.int32), enc_lens=elens.numpy().astype(np.int32),
                        alphas=alphas.numpy(), peaks=peaks.numpy(), acoustic=emb[:, :, ::5].numpy(),
                        us_alphas=us_alphas.numpy(), us_peaks=us_peaks.numpy(),
                        logp_rows=np.array(keep, dtype=np.int32), logp_sel=logp[:, keep, :].numpy(),
                        ids_flat=np.array([t for r in ids for t in r], dtype=np.int32), ids_len=np.array([len(r) for r in ids], dtype=np.int32),
                        stamps_flat=np.array([v for st in stamps for pair in st for v in pair], dtype=np.int32),
                        stamps_len=np.array([len(st) for st in stamps], dtype=np.int32))
    valid = torch.arange(N)[None, :] < tokl[:, None]
    print("%s: B=%d tokens=%s stamps=%s min margin %.3e first stamps %s" % (name, len(wavs), tokl.tolist(), [len(s) for s in stamps],
                                                                           float((top2[..., 0] - top2[..., 1])[valid].min()), stamps[0][:3]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bicif":
    ref_shim.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        for name, (cfg, wseed, specs) in BICIF_CASES.items():
            run_bicif_case(name, cfg, wseed, specs, tmp)


# ------------------------------------------------------------------------------------------------ SeacoParaformer
SEACO_CASES = {
    # name: (cfg, weight seed, wavs, n hotwords, nfilter)  — the second case exercises attention-score filtering (nfilter < hotwords)
    "seaco_tiny_ragged3": (synth.PARAFORMER_TINY, 10, [(48000, 41, "speechlike"), (27200, 42, "noise"), (38437, 43, "speechlike")], 6, 50),
    "seaco_tiny_asf": (synth.PARAFORMER_TINY, 11, [(40000, 44, "speechlike"), (30000, 45, "speechlike")], 24, 8),
}


def run_seaco_case(name, cfg, wseed, wav_specs, n_hot, nfilter, tmp):
    from funasr import AutoModel
    cmvn_file = os.path.join(tmp, "am_%s.mvn" % name)
    write_cmvn_file(cmvn_file, synth.make_cmvn(cfg, seed=1))
    pt = os.path.join(tmp, "seaco_%s.pt" % name)
    torch.save(synth.make_seaco_state_dict(cfg, wseed), pt)
    tokens = ["<blank>", "<s>", "</s>"] + ["t%d" % i for i in range(cfg.vocab - 4)] + ["<unk>"]
    no_bias = synth.seaco_no_bias_id(cfg)
    am = AutoModel(
        model="SeacoParaformer",
        model_conf=dict(ctc_weight=0.0, lsm_weight=0.1, length_normalized_loss=True, predictor_weight=1.0, predictor_bias=1, sampling_ratio=0.75,
                        inner_dim=512, bias_encoder_type="lstm", bias_encoder_bid=False, seaco_lsm_weight=0.1, seaco_length_normal=True,
                        train_decoder=False, NO_BIAS=no_bias),
        encoder="SANMEncoder",
        encoder_conf=dict(output_size=cfg.d_model, attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.enc_layers,
                          dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1, input_layer="pe",
                          pos_enc_class="SinusoidalPositionEncoder", normalize_before=True, kernel_size=cfg.kernel, sanm_shfit=0,
                          selfattention_layer_type="sanm"),
        decoder="ParaformerSANMDecoder",
        decoder_conf=dict(attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.dec_layers, dropout_rate=0.1,
                          positional_dropout_rate=0.1, self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1,
                          att_layer_num=cfg.dec_layers, kernel_size=cfg.kernel, sanm_shfit=0),
        seaco_decoder="ParaformerSANMDecoder",
        seaco_decoder_conf=dict(attention_heads=4, linear_units=synth.SEACO_FFN, num_blocks=4, dropout_rate=0.1, positional_dropout_rate=0.1,
                                self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1, kernel_size=synth.SEACO_KERNEL, sanm_shfit=0,
                                use_output_layer=False, wo_input_layer=True),
        predictor="CifPredictorV3",
        predictor_conf=dict(idim=cfg.d_model, threshold=1.0, l_order=1, r_order=1, tail_threshold=cfg.tail_threshold, smooth_factor2=0.25,
                            noise_threshold2=0.01, upsample_times=3, use_cif1_cnn=False, upsample_type="cnn_blstm"),
        frontend="WavFrontend",
        frontend_conf=dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6, dither=0.0,
                           cmvn_file=cmvn_file),
        tokenizer="CharTokenizer", tokenizer_conf=dict(token_list=tokens, unk_symbol="<unk>", split_with_space=True),
        device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt,
    )
    model, frontend = am.model, am.kwargs["frontend"]
    hw_list = synth.make_hotwords(n_hot, cfg.vocab, seed=9)
    wavs = [synth.make_wav(n, s, k) for (n, s, k) in wav_specs]
    with torch.no_grad():
        from funasr.utils.load_utils import extract_fbank
        feats, flens = extract_fbank([w for w in wavs], frontend=frontend)
        enc, elens = model.encode(feats, flens)
        emb, tok, alphas, peaks = model.calc_predictor(enc, elens)
        tokl = tok.round().long()
        merged = model._seaco_decode_with_ASF(enc, elens, emb, tokl, hw_list=hw_list, nfilter=nfilter)
        sel = model._hotword_representation(torch.nn.utils.rnn.pad_sequence([torch.tensor(h) for h in hw_list], batch_first=True),
                                            torch.tensor([len(h) for h in hw_list]).int())
    ids = []
    for i in range(len(wavs)):
        ys = merged[i, : int(tokl[i])].argmax(-1).tolist()
        ids.append([t for t in ys if t not in (0, 1, 2)])
    N = merged.shape[1]
    keep = sorted(set(list(range(min(4, N))) + [N // 2, N - 1]))
    top2 = torch.topk(merged, 2, dim=-1).values
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), token_num=tokl.numpy().astype(np.int32), hw_selected=sel.numpy(),
                        logp_rows=np.array(keep, dtype=np.int32), merged_sel=merged[:, keep, :].numpy(),
                        argmax=merged.argmax(-1).numpy().astype(np.int32), margin=(top2[..., 0] - top2[..., 1]).numpy(),
                        ids_flat=np.array([t for r in ids for t in r], dtype=np.int32), ids_len=np.array([len(r) for r in ids], dtype=np.int32))
    valid = torch.arange(N)[None, :] < tokl[:, None]
    n_bias = int(((merged.argmax(-1) != 0) & valid).sum())
    print("%s: B=%d tokens=%s n_hot=%d nfilter=%d min margin %.3e ids[0][:6]=%s" % (name, len(wavs), tokl.tolist(), len(hw_list), nfilter,
                                                                                 float((top2[..., 0] - top2[..., 1])[valid].min()), ids[0][:6]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "seaco":
    ref_shim.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        for name, (cfg, wseed, specs, n_hot, nfilter) in SEACO_CASES.items():
            run_seaco_case(name, cfg, wseed, specs, n_hot, nfilter, tmp)
