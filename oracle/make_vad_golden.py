"""Generate tests/golden/vad_*.npz / longaudio_*.npz by running the UNMODIFIED reference (FsmnVADStreaming + WavFrontendOnline +
AutoModel.inference_with_vad) on CPU.  Run in the build container only:   python oracle/make_vad_golden.py

The per-frame scores / frame energies the reference computed are recorded by wrapping two bound methods of the model INSTANCE in
this script (the reference's code is untouched); segments come from AutoModel.generate()."""
import os
import re
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_runner  # noqa: E402
import ref_shim  # noqa: E402
from funasr_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name: (seconds, wav seed, pattern [(speech_s, silence_s)...] or None, generate kwargs)
VAD_CASES = {
    "vad_30s": (30.0, 1, [(3.0, 2.5), (1.5, 0.4), (4.0, 3.0), (2.0, 2.2)], {}),
    "vad_130s": (130.0, 2, [(70.0, 2.5), (5.0, 0.3), (20.0, 2.1), (10.0, 3.0)], {}),
    "vad_fixed800": (30.0, 3, [(2.0, 1.0), (3.0, 0.5), (1.0, 1.5)], {"max_end_silence_time": 800}),
    "vad_random45": (45.0, 4, None, {}),
    "vad_short": (1.2, 5, [(5.0, 0.1)], {}),
    "vad_silence": (3.0, 6, [(0.0, 9.0)], {}),
}
VAD_WEIGHT_SEED = 0


def vad_conf(cmvn_file):
    c = synth.VAD_DEFAULT
    return dict(
        model="FsmnVADStreaming",
        model_conf=dict(sample_rate=16000, detect_mode=1, snr_mode=0, max_end_silence_time=800, max_start_silence_time=3000,
                        do_start_point_detection=True, do_end_point_detection=True, window_size_ms=200, sil_to_speech_time_thres=150,
                        speech_to_sil_time_thres=150, speech_2_noise_ratio=1.0, do_extend=1, lookback_time_start_point=200,
                        lookahead_time_end_point=100, max_single_segment_time=60000, snr_thres=-100.0, noise_frame_num_used_for_snr=100,
                        decibel_thres=-100.0, speech_noise_thres=0.6, fe_prior_thres=0.0001, silence_pdf_num=1, sil_pdf_ids=[0],
                        speech_noise_thresh_low=-0.1, speech_noise_thresh_high=0.3, output_frame_probs=False, frame_in_ms=10,
                        frame_length_ms=25),
        encoder="FSMN",
        encoder_conf=dict(input_dim=c.input_dim, input_affine_dim=c.input_affine_dim, fsmn_layers=c.fsmn_layers, linear_dim=c.linear_dim,
                          proj_dim=c.proj_dim, lorder=c.lorder, rorder=c.rorder, lstride=1, rstride=0, output_affine_dim=c.output_affine_dim,
                          output_dim=c.output_dim),
        frontend="WavFrontendOnline",
        frontend_conf=dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, dither=0.0, lfr_m=5, lfr_n=1,
                           cmvn_file=cmvn_file))


def build_vad(tmp, device="cpu"):
    from funasr import AutoModel
    cmvn_file = os.path.join(tmp, "vad.mvn")
    ref_runner.write_cmvn_file(cmvn_file, synth.make_vad_cmvn(0))
    pt = os.path.join(tmp, "vad.pt")
    torch.save(synth.make_vad_state_dict(synth.VAD_DEFAULT, VAD_WEIGHT_SEED), pt)
    return AutoModel(**vad_conf(cmvn_file), device=device, ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt)


def run_vad_case(am, name, seconds, seed, pattern, gen_kw):
    wav = synth.make_vad_wav(seconds, seed, pattern)
    model = am.model
    rec = {"scores": [], "db": []}
    enc_fwd = model.encoder.forward
    comp_db = model.ComputeDecibel

    def enc_wrapped(feats, cache=None):
        out = enc_fwd(feats, cache=cache)
        rec["scores"].append(out.detach().clone())
        return out

    def db_wrapped(cache=None, frame_count=None):
        before = len(cache["stats"].decibel)
        comp_db(cache=cache, frame_count=frame_count)
        rec["db"].extend(cache["stats"].decibel[before:])

    model.encoder.forward = enc_wrapped
    model.ComputeDecibel = db_wrapped
    try:
        res = am.generate(input=wav.numpy(), disable_pbar=True, **gen_kw)
    finally:
        model.encoder.forward = enc_fwd
        model.ComputeDecibel = comp_db
    segs = res[0]["value"]
    scores = torch.cat(rec["scores"], dim=1)[0] if rec["scores"] else torch.zeros(0, 248)
    rows = sorted(set([0, 1, scores.shape[0] // 2, max(scores.shape[0] - 1, 0)])) if scores.shape[0] else []
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), segments=np.array(segs, dtype=np.int64).reshape(-1, 2),
                        sil_prob=scores[:, 0].numpy(), decibel=np.array(rec["db"], dtype=np.float64),
                        chunk_frames=np.array([s.shape[1] for s in rec["scores"]], dtype=np.int64),
                        score_rows=np.array(rows, dtype=np.int64), score_sel=scores[rows].numpy() if rows else np.zeros((0, 248), np.float32),
                        n_samples=np.int64(wav.numel()))
    print("%s: %.1f s -> %d frames in chunks %s, %d segments %s" % (name, seconds, scores.shape[0], [s.shape[1] for s in rec["scores"]],
                                                                  len(segs), segs[:6]))


# ---- long-audio ASR: AutoModel(model=Paraformer, vad_model=FsmnVADStreaming).generate() -> inference_with_vad
LONG_CASES = {
    # name: (seconds, wav seed, pattern, generate kwargs)
    "longaudio_40s": (40.0, 7, [(3.0, 2.5), (1.5, 2.2), (4.0, 3.0), (2.0, 2.2), (6.0, 2.4)], {"batch_size_s": 6, "pred_timestamp": True}),
    "longaudio_25s_onebatch": (25.0, 8, [(2.0, 2.5), (3.0, 2.1)], {"batch_size_s": 300}),
}


def run_long_case(name, seconds, seed, pattern, gen_kw, tmp):
    from funasr import AutoModel
    cfg = synth.PARAFORMER_TINY
    asr_cmvn = os.path.join(tmp, "asr.mvn")
    ref_runner.write_cmvn_file(asr_cmvn, synth.make_cmvn(cfg, 1))
    vad_cmvn = os.path.join(tmp, "vad2.mvn")
    ref_runner.write_cmvn_file(vad_cmvn, synth.make_vad_cmvn(0))
    pt = os.path.join(tmp, "asr.pt")
    torch.save(synth.make_state_dict(cfg, 3), pt)
    vpt = os.path.join(tmp, "vad2.pt")
    torch.save(synth.make_vad_state_dict(synth.VAD_DEFAULT, VAD_WEIGHT_SEED), vpt)
    vc = vad_conf(vad_cmvn)
    am = AutoModel(model="Paraformer",
                   model_conf=dict(ctc_weight=0.0, lsm_weight=0.1, length_normalized_loss=True, predictor_weight=1.0, predictor_bias=1, sampling_ratio=0.75),
                   encoder="SANMEncoder", encoder_conf=ref_runner._enc_conf(cfg), decoder="ParaformerSANMDecoder", decoder_conf=ref_runner._dec_conf(cfg),
                   predictor="CifPredictorV2", predictor_conf=dict(idim=cfg.d_model, threshold=1.0, l_order=1, r_order=1, tail_threshold=cfg.tail_threshold),
                   frontend="WavFrontend", frontend_conf=ref_runner._frontend_conf(asr_cmvn), tokenizer="CharTokenizer",
                   tokenizer_conf=dict(token_list=ref_runner.token_list(cfg), unk_symbol="<unk>", split_with_space=True),
                   init_param=pt, vad_model=vc["model"],
                   vad_kwargs=dict(model_conf=vc["model_conf"], encoder=vc["encoder"], encoder_conf=vc["encoder_conf"], frontend=vc["frontend"],
                                   frontend_conf=vc["frontend_conf"], init_param=vpt),
                   device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True)
    wav = synth.make_vad_wav(seconds, seed, pattern)
    # the STRING "cpu" forces batch_size 0 in inference_with_vad (auto_model.py:929-930: one segment per call); a torch.device
    # object does not compare equal to it, so the reference's dynamic batching over the duration-sorted segments runs — on the CPU
    res = am.generate(input=wav.numpy(), disable_pbar=True, device=torch.device("cpu"), **gen_kw)
    r = res[0]
    text = r.get("text", "")
    ids = [int(x) + 3 for x in re.findall(r"t(\d+)", text)]
    ts = r.get("timestamp", [])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), ids=np.array(ids, dtype=np.int64), timestamp=np.array(ts, dtype=np.int64).reshape(-1, 2),
                        n_samples=np.int64(wav.numel()), text=np.array(text))
    print("%s: %d ids, %d stamps, text[:60]=%r" % (name, len(ids), len(ts), text[:60]))


# ---- CT-Transformer punctuation: AutoModel(model="CTTransformer").generate(input=text)
PUNC_CASES = {"punc_short": (12, 1), "punc_long": (140, 2), "punc_english_tail": (33, 3)}


def run_punc_cases(tmp):
    from funasr import AutoModel
    pt = os.path.join(tmp, "punc.pt")
    torch.save(synth.make_punc_state_dict(0), pt)
    toks = synth.punc_token_list()
    am = AutoModel(model="CTTransformer",
                   model_conf=dict(ignore_id=0, embed_unit=synth.PUNC_DIM, att_unit=synth.PUNC_DIM, dropout_rate=0.1, punc_list=synth.PUNC_LIST,
                                   punc_weight=[1.0] * len(synth.PUNC_LIST), sentence_end_id=3),
                   encoder="SANMEncoder",
                   encoder_conf=dict(input_size=synth.PUNC_DIM, output_size=synth.PUNC_DIM, attention_heads=synth.PUNC_HEADS, linear_units=synth.PUNC_FFN,
                                     num_blocks=synth.PUNC_LAYERS, dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.0,
                                     input_layer="pe", pos_enc_class="SinusoidalPositionEncoder", normalize_before=True, kernel_size=11, sanm_shfit=0,
                                     selfattention_layer_type="sanm", padding_idx=0),
                   tokenizer="CharTokenizer", tokenizer_conf=dict(token_list=toks, unk_symbol="<unk>"),
                   device="cpu", ncpu=os.cpu_count(), disable_update=True, disable_pbar=True, init_param=pt)
    for name, (n_words, seed) in PUNC_CASES.items():
        text = synth.make_punc_text(n_words, seed)
        res = am.generate(input=text, disable_pbar=True)
        r = res[0]
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), text_in=np.array(text), text_out=np.array(r["text"]),
                            punc_array=np.asarray(r["punc_array"]).astype(np.int64))
        print("%s: %d words -> %r  punc %s" % (name, n_words, r["text"][:70], np.bincount(np.asarray(r["punc_array"]).astype(np.int64), minlength=6).tolist()))


def main():
    ref_shim.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["vad", "long", "punc"]
    with tempfile.TemporaryDirectory() as tmp:
        if "vad" in which:
            am = build_vad(tmp)
            for name, (seconds, seed, pattern, kw) in VAD_CASES.items():
                run_vad_case(am, name, seconds, seed, pattern, kw)
        if "long" in which:
            for name, (seconds, seed, pattern, kw) in LONG_CASES.items():
                run_long_case(name, seconds, seed, pattern, kw, tmp)
        if "punc" in which:
            run_punc_cases(tmp)


if __name__ == "__main__":
    main()
