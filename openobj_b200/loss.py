"""Mirror of objnerf/loss.py: step_batch_loss with the reference's signature, backed by the fused K3 kernels
(differentiable w.r.t. alpha, color and pred_partfeat through torch.autograd.Function)."""
from . import ops

last_terms = None     # per-object [N,4] (depth, colour, opacity, feature) of the most recent call
last_flags = None     # OO_FLAG_* bits of the most recent call (device int32[1]); bit 0 = the reference would exit(-1)


def step_batch_loss(alpha, color, gt_depth, gt_color, sem_labels, mask_depth, z_vals,
                    color_scaling=5.0, opacity_scaling=10.0, gt_partfeat=None, pred_partfeat=None, partfeat_scaling=5.0):
    """loss.py:5-103.  `mask_depth` is accepted and ignored exactly like the reference."""
    global last_terms, last_flags
    loss, last_terms, last_flags = ops.step_loss(alpha, color, gt_depth, gt_color, sem_labels, z_vals,
                                                 pred_feat=pred_partfeat if gt_partfeat is not None else None,
                                                 gt_feat=gt_partfeat, color_scaling=color_scaling,
                                                 opacity_scaling=opacity_scaling, feat_scaling=partfeat_scaling)
    return loss, None
