"""Generate the golden fixtures from the UNMODIFIED reference (needs /root/reference; run here, not on
the GPU box):   python tests/golden/make_golden.py
Outputs (committed): golden_host.json, golden_masks.pt, golden_step_c1.pt
"""
import copy
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, HERE)
from common import C1, VITL16_MASKS, sha16, synth_clips, synth_state  # noqa: E402

from src.models.utils.pos_embs import get_3d_sincos_pos_embed  # noqa: E402  (reference)
from src.masks.multiblock3d import MaskCollator as RefMB3D  # noqa: E402
from src.masks.random_tube import MaskCollator as RefTube  # noqa: E402
from src.masks.utils import apply_masks as ref_apply_masks  # noqa: E402
from src.utils.tensors import repeat_interleave_batch as ref_rib, trunc_normal_ as ref_trunc  # noqa: E402
from src.utils.schedulers import WarmupCosineSchedule, CosineWDSchedule  # noqa: E402
import src.models.vision_transformer as ref_vit  # noqa: E402
import src.models.predictor as ref_pred  # noqa: E402
from src.models.utils.multimask import MultiMaskWrapper, PredictorMultiMaskWrapper  # noqa: E402

torch.set_num_threads(8)
host = {}

# ---- positional tables -------------------------------------------------------------------------
host['pos_embed'] = {}
for (D, grid, depth) in [(1024, 14, 8), (1280, 14, 8), (1280, 24, 8), (384, 14, 8), (384, 24, 8), (192, 14, 4)]:
    e = get_3d_sincos_pos_embed(D, grid, depth, cls_token=False, uniform_power=True)
    host['pos_embed'][f'{D}_{grid}_{depth}'] = dict(sum=float(e.sum()), sha=sha16(torch.from_numpy(e).float()))

# ---- masks ------------------------------------------------------------------------------------
masks_pt = {}
host['masks'] = {}
for tag, (crop, T, B) in dict(c2=(224, 16, 32), c4=(224, 16, 24), c5=(384, 16, 10), c1=(224, 8, 2)).items():
    torch.manual_seed(0)
    coll = RefMB3D(cfgs_mask=VITL16_MASKS, crop_size=crop, num_frames=T, patch_size=16, tubelet_size=2)
    entry = []
    for call in range(2):
        _, me, mp = coll([torch.zeros(1) for _ in range(B)])
        h = hashlib.sha256()
        for t in me + mp:
            h.update(t.numpy().tobytes())
        entry.append(dict(Ke=[int(t.shape[1]) for t in me], Kp=[int(t.shape[1]) for t in mp], sha=h.hexdigest()[:16],
                          first20=me[0][0][:20].tolist()))
        if tag == 'c1':
            masks_pt[f'c1_call{call}'] = dict(enc=me, pred=mp)
    host['masks'][tag] = entry
np.random.seed(0)
tube = RefTube(cfgs_mask=[dict(ratio=0.9)], crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
_, te, tp = tube([torch.zeros(1) for _ in range(4)])
host['random_tube'] = dict(Ke=int(te[0].shape[1]), Kp=int(tp[0].shape[1]), sha_enc=sha16(te[0]), sha_pred=sha16(tp[0]))

# ---- small helpers ------------------------------------------------------------------------------
host['repeat_interleave'] = ref_rib(torch.arange(6), 2, 2).tolist()
t = torch.empty(64, 32)
torch.manual_seed(3)
ref_trunc(t, std=0.02)
host['trunc_normal'] = dict(sha=sha16(t), std=float(t.std()))


class _FakeOpt:
    def __init__(self):
        self.param_groups = [dict(lr=0., weight_decay=0.), dict(lr=0., weight_decay=0., WD_exclude=True)]


opt = _FakeOpt()
sch = WarmupCosineSchedule(opt, warmup_steps=12, start_lr=0.0002, ref_lr=0.000625, final_lr=1e-6, T_max=100)
wds = CosineWDSchedule(opt, ref_wd=0.04, final_wd=0.4, T_max=100)
host['schedules'] = dict(lr=[sch.step() for _ in range(100)], wd=[wds.step() for _ in range(100)],
                         wd_excluded=opt.param_groups[1]['weight_decay'])

# ---- one full C1 step through the reference modules (fp32, CPU) --------------------------------------
cfg = C1
enc = ref_vit.__dict__[cfg['model_name']](img_size=cfg['crop_size'], patch_size=cfg['patch_size'],
                                           num_frames=cfg['num_frames'], tubelet_size=cfg['tubelet_size'],
                                           uniform_power=True, use_sdpa=True)
enc = MultiMaskWrapper(enc)
pred = ref_pred.__dict__['vit_predictor'](img_size=cfg['crop_size'], use_mask_tokens=True, patch_size=cfg['patch_size'],
                                           num_frames=cfg['num_frames'], tubelet_size=cfg['tubelet_size'],
                                           embed_dim=enc.backbone.embed_dim, predictor_embed_dim=cfg['pred_embed_dim'],
                                           depth=cfg['pred_depth'], num_heads=enc.backbone.num_heads, uniform_power=True,
                                           num_mask_tokens=2, zero_init_mask_tokens=True, use_sdpa=True)
pred = PredictorMultiMaskWrapper(pred)
enc_shapes = {k: tuple(v.shape) for k, v in enc.backbone.state_dict().items()}
pred_shapes = {k: tuple(v.shape) for k, v in pred.backbone.state_dict().items()}
host['state_shapes'] = dict(encoder={k: list(v) for k, v in enc_shapes.items()},
                            predictor={k: list(v) for k, v in pred_shapes.items()})
host['param_order'] = dict(encoder=[n for n, _ in enc.backbone.named_parameters()],
                           predictor=[n for n, _ in pred.backbone.named_parameters()])
enc.backbone.load_state_dict(synth_state(enc_shapes, seed=11, keep=('pos_embed',)), strict=False)
pred.backbone.load_state_dict(synth_state(pred_shapes, seed=12, keep=('pos_embed',)), strict=False)
tgt = copy.deepcopy(enc)
tgt.backbone.load_state_dict(synth_state(enc_shapes, seed=13, keep=('pos_embed',)), strict=False)
for p in tgt.parameters():
    p.requires_grad = False

clips = synth_clips(cfg['batch'], cfg['num_frames'], cfg['crop_size'], cfg['crop_size'], seed=0)
me, mp = masks_pt['c1_call0']['enc'], masks_pt['c1_call0']['pred']
import torch.nn.functional as F  # noqa: E402

with torch.no_grad():
    h = tgt(clips)
    h = F.layer_norm(h, (h.size(-1),))
    h = ref_apply_masks(h, mp, concat=False)
z_enc = enc(clips, me)
z = pred(z_enc, h, me, mp)
loss_jepa = sum(torch.mean(torch.abs(zi - hi)) for zi, hi in zip(z, h)) / len(mp)
pstd = sum(torch.sqrt(zi.var(dim=1) + 0.0001) for zi in z) / len(z)
loss_reg = torch.mean(F.relu(1. - pstd))
loss_jepa.backward()

step = dict(loss_jepa=float(loss_jepa), loss_reg=float(loss_reg))
step['h_slices'] = [hi[:, :4, :16].clone() for hi in h]
step['z_slices'] = [zi[:, :4, :16].detach().clone() for zi in z]
step['zenc_slices'] = [zi[:, :4, :16].detach().clone() for zi in z_enc]
step['h_norm'] = [float(hi.norm()) for hi in h]
step['z_norm'] = [float(zi.norm()) for zi in z]
step['zenc_norm'] = [float(zi.norm()) for zi in z_enc]
step['enc_grad_norm'] = {n: float(p.grad.norm()) for n, p in enc.backbone.named_parameters() if p.grad is not None}
step['pred_grad_norm'] = {n: float(p.grad.norm()) for n, p in pred.backbone.named_parameters() if p.grad is not None}
step['enc_grad_slices'] = {n: p.grad.reshape(-1)[:32].clone() for n, p in enc.backbone.named_parameters()
                           if p.grad is not None and ('blocks.0.' in n or 'blocks.11.' in n or 'patch_embed' in n or n.startswith('norm'))}
step['pred_grad_slices'] = {n: p.grad.reshape(-1)[:32].clone() for n, p in pred.backbone.named_parameters()
                            if p.grad is not None and ('blocks.0.' in n or 'blocks.11.' in n or 'predictor_embed' in n
                                                       or 'predictor_proj' in n or 'mask_tokens' in n or 'predictor_norm' in n)}
# EMA (train.py:484-487)
m = 0.998
with torch.no_grad():
    for pq, pk in zip(enc.parameters(), tgt.parameters()):
        pk.data.mul_(m).add_((1. - m) * pq.detach().data)
step['ema_slices'] = {n: p.data.reshape(-1)[:32].clone() for n, p in tgt.backbone.named_parameters()
                      if 'blocks.3.' in n or n == 'pos_embed'}

torch.save(masks_pt, os.path.join(HERE, 'golden_masks.pt'))
torch.save(step, os.path.join(HERE, 'golden_step_c1.pt'))
with open(os.path.join(HERE, 'golden_host.json'), 'w') as f:
    json.dump(host, f, indent=1)
print('loss_jepa', step['loss_jepa'], 'loss_reg', step['loss_reg'])
print('wrote fixtures to', HERE)
