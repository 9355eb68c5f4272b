"""Model / optimizer / checkpoint factories with the reference's signatures (app/vjepa/utils.py:28-210)."""
import logging
import sys

import torch

import src.models.predictor as vit_pred
import src.models.vision_transformer as video_vit
from jepa_b200.optim import FlatAdamW, FlatGradScaler
from src.models.utils.multimask import MultiMaskWrapper, PredictorMultiMaskWrapper
from src.utils.schedulers import CosineWDSchedule, WarmupCosineSchedule
from src.utils.tensors import trunc_normal_

logging.basicConfig(stream=sys.stdout, level=logging.INFO)
logger = logging.getLogger()


def load_checkpoint(r_path, encoder, predictor, target_encoder, opt, scaler):
    """Restore encoder / predictor / target / optimizer / scaler from a reference-format .pth.tar.
    Any failure is logged and training restarts from epoch 0, as in the reference (utils.py:28-83)."""
    try:
        checkpoint = torch.load(r_path, map_location=torch.device('cpu'))
    except Exception as e:
        logger.info(f'Encountered exception when loading checkpoint {e}')
        return encoder, predictor, target_encoder, opt, scaler, 0
    epoch = 0
    try:
        epoch = checkpoint['epoch']
        for tag, net in (('encoder', encoder), ('predictor', predictor), ('target_encoder', target_encoder)):
            if net is None:
                continue
            msg = net.load_state_dict(checkpoint[tag])
            logger.info(f'loaded pretrained {tag} from epoch {epoch} with msg: {msg}')
        opt.load_state_dict(checkpoint['opt'])
        if scaler is not None:
            scaler.load_state_dict(checkpoint['scaler'])
        logger.info(f'loaded optimizers from epoch {epoch}')
        logger.info(f'read-path: {r_path}')
        del checkpoint
    except Exception as e:
        logger.info(f'Encountered exception when loading checkpoint {e}')
        epoch = 0
    return encoder, predictor, target_encoder, opt, scaler, epoch


def init_video_model(device, patch_size=16, num_frames=16, tubelet_size=2, model_name='vit_base', crop_size=224,
                     pred_depth=6, pred_embed_dim=384, uniform_power=False, use_mask_tokens=False, num_mask_tokens=2,
                     zero_init_mask_tokens=True, use_sdpa=False):
    """Build (MultiMaskWrapper(encoder), PredictorMultiMaskWrapper(predictor)); note the predictor inherits the
    ENCODER's head count (utils.py:119) and every Linear / LayerNorm is re-initialised after construction, which
    undoes the per-layer rescale of proj / fc2 (utils.py:127-140)."""
    encoder = video_vit.__dict__[model_name](img_size=crop_size, patch_size=patch_size, num_frames=num_frames,
                                             tubelet_size=tubelet_size, uniform_power=uniform_power, use_sdpa=use_sdpa)
    encoder = MultiMaskWrapper(encoder)
    predictor = vit_pred.__dict__['vit_predictor'](
        img_size=crop_size, use_mask_tokens=use_mask_tokens, patch_size=patch_size, num_frames=num_frames,
        tubelet_size=tubelet_size, embed_dim=encoder.backbone.embed_dim, predictor_embed_dim=pred_embed_dim,
        depth=pred_depth, num_heads=encoder.backbone.num_heads, uniform_power=uniform_power,
        num_mask_tokens=num_mask_tokens, zero_init_mask_tokens=zero_init_mask_tokens, use_sdpa=use_sdpa)
    predictor = PredictorMultiMaskWrapper(predictor)

    def reinit(m):
        if isinstance(m, torch.nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)
        elif isinstance(m, torch.nn.LayerNorm):
            torch.nn.init.constant_(m.bias, 0)
            torch.nn.init.constant_(m.weight, 1.0)

    for net in (encoder, predictor):
        for m in net.modules():
            reinit(m)
        net.to(device)
    logger.info(encoder)
    logger.info(predictor)

    def count_parameters(model):
        return sum(p.numel() for p in model.parameters() if p.requires_grad)

    logger.info(f'Encoder number of parameters: {count_parameters(encoder)}')
    logger.info(f'Predictor number of parameters: {count_parameters(predictor)}')
    return encoder, predictor


def init_opt(encoder, predictor, iterations_per_epoch, start_lr, ref_lr, warmup, num_epochs, wd=1e-6, final_wd=1e-6,
             final_lr=0.0, mixed_precision=False, ipe_scale=1.25, betas=(0.9, 0.999), eps=1e-8, zero_init_bias_wd=True):
    """Four AdamW groups (enc weights, pred weights, enc bias/1-D, pred bias/1-D; the latter two excluded from the
    weight-decay schedule), warm-up-cosine LR, cosine WD, GradScaler when mixed precision (utils.py:156-210)."""
    def is_nodecay(n, p):
        return ('bias' in n) or (len(p.shape) == 1)

    param_groups = [
        {'params': [p for n, p in encoder.named_parameters() if not is_nodecay(n, p)]},
        {'params': [p for n, p in predictor.named_parameters() if not is_nodecay(n, p)]},
        {'params': [p for n, p in encoder.named_parameters() if is_nodecay(n, p)],
         'WD_exclude': zero_init_bias_wd, 'weight_decay': 0},
        {'params': [p for n, p in predictor.named_parameters() if is_nodecay(n, p)],
         'WD_exclude': zero_init_bias_wd, 'weight_decay': 0},
    ]
    logger.info('Using AdamW')
    optimizer = FlatAdamW(param_groups, betas=betas, eps=eps)
    total = int(ipe_scale * num_epochs * iterations_per_epoch)
    scheduler = WarmupCosineSchedule(optimizer, warmup_steps=int(warmup * iterations_per_epoch), start_lr=start_lr,
                                     ref_lr=ref_lr, final_lr=final_lr, T_max=total)
    wd_scheduler = CosineWDSchedule(optimizer, ref_wd=wd, final_wd=final_wd, T_max=total)
    scaler = FlatGradScaler() if mixed_precision else None   # torch.cuda.amp.GradScaler with a fused flat-buffer unscale
    return optimizer, scaler, scheduler, wd_scheduler
