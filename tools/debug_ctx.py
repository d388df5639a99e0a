import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from radtts_b200 import configs, synth, ops
from radtts_b200.radtts import RADTTS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = np.load('tests/golden/radtts_forward.npz')
torch.manual_seed(0)
m = RADTTS(**configs.model_config("radtts")).eval(); synth.load_synth(m, 1234); m = m.cuda()
b = {k: v.cuda() for k, v in synth.synth_batch(3,70,24,seed=1234).items()}
ops.set_precision("fp32")
with torch.no_grad():
    spk = m.encode_speaker(b['speaker_ids'])
    text_enc, emb = m.encode_text(b['text'], b['in_lens'])
    attn = torch.from_numpy(g['attn']).cuda()
    context = torch.bmm(text_enc, attn.squeeze(1).transpose(1,2))
    ctx = m.preprocess_context(context, spk, b['out_lens'], None, None)
    ref = torch.from_numpy(g['context']).cuda()
    lens = b['out_lens']//2
    for i in range(3):
        print('ctx diff', i, (ctx[i,:,:lens[i]]-ref[i,:,:lens[i]]).abs().max().item())
    z1, _, _ = ops.decoder_forward(m, b['mel'], ctx, b['out_lens'])
    z2, _, _ = ops.decoder_forward(m, b['mel'], ref, b['out_lens'])
    zr = torch.from_numpy(g['z_mel']).cuda()
    for i in range(3):
        print('z(ctx gpu) vs ref', i, (z1[i,:,:lens[i]]-zr[i,:,:lens[i]]).abs().max().item(), ' z(ctx gold) vs ref', (z2[i,:,:lens[i]]-zr[i,:,:lens[i]]).abs().max().item())
    out = m(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True, attn_prior=b["attn_prior"])
    for i in range(3):
        print('full fwd z vs ref', i, (out['z_mel'][i,:,:lens[i]]-zr[i,:,:lens[i]]).abs().max().item(), 'attn diff', (out['attn'][i]-attn[i]).abs().max().item())
