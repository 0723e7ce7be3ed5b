import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import load_sv_case
from funasr_b200 import synth
from funasr_b200.engine import SenseVoiceEngine
DEV = torch.device('cuda:0')
name = 'sv_large_single'
cfg, wseed, wavs, cmvn, g = load_sv_case(name)
res = {}
for mode in ['fp32', 'fp16x3']:
    eng = SenseVoiceEngine(synth.make_sensevoice_state_dict(cfg, wseed), cfg, DEV, gemm_mode=mode, cmvn=cmvn)
    lens = [w.numel() for w in wavs]
    pad = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True).to(DEV)
    o = eng.forward_wav(pad, torch.tensor(lens, dtype=torch.int32, device=DEV), lens, language_id=0, textnorm_id=15, want_taps=True)
    torch.cuda.synchronize()
    res[mode] = o
    am = o['argmax'].cpu().numpy(); ga = g['argmax']
    valid = np.arange(ga.shape[1])[None, :] < g['enc_lens'][:, None]
    bad = np.argwhere((am != ga) & valid)
    print(mode, 'T', ga.shape, 'mismatches', len(bad), bad[:10].tolist())
    lp = o['logp'][0].float().cpu()
    for b_, t_ in bad[:10]:
        top = torch.topk(lp[t_], 3)
        print('  t', t_, 'ours', am[b_, t_], 'gold', ga[b_, t_], 'top3', top.indices.tolist(), [f'{x:.6f}' for x in top.values.tolist()], 'gold logp ours', float(lp[t_, ga[b_, t_]]))
d = (res['fp32']['logp'] - res['fp16x3']['logp']).abs().max()
print('max |logp fp32 - fp16x3|', float(d))
