"""Drop-in for the reference's ``src/utils/proposal_utils.py`` (same names, arguments, defaults and
return types).  Each call is one fused selection kernel (+ three NMS launches for inference) instead
of the reference's sigmoid / permute copy / topk / mask-index / gather / clamp / mask-index chain;
the only host sync left is the one the dense return type demands (reading the count)."""
import torch

from ... import ops


def sample_proposals(proposals, num_samples=128):
    """Randomly sample proposals for training (proposal_utils.py:6-10).  Stays on torch's device RNG
    so that the random stream matches the reference's."""
    num_samples = min(num_samples, len(proposals))
    sampled_indices = torch.randperm(len(proposals), device=proposals.device)[:num_samples]
    return proposals[sampled_indices], sampled_indices


def _select(cls_scores, anchors, image_size, k, score_threshold, min_box_size):
    obj = cls_scores.unsqueeze(0)
    boxes, scores, index, counts = ops.rpn_select(
        [obj], k=min(int(k), obj[0].numel()), img_size=image_size, score_thresh=score_threshold, min_size=min_box_size,
        anchors_per_level=[anchors], score_strict=True, topk_on_sigmoid=True)
    return boxes[0], scores[0], counts[0]          # [1,k,4], [1,k], [1]


def generate_training_proposals(cls_scores, anchors, image_size, device,
                                num_proposals=500, score_threshold=0.01,
                                min_box_size=5):
    """Proposals for training: top-k on sigmoid scores, score > thr, clip, min-size; no NMS
    (proposal_utils.py:12-31).  cls_scores [A,h,w], anchors [h*w*A,4] -> f32[K,4]."""
    boxes, _, counts = _select(cls_scores, anchors, image_size, num_proposals, score_threshold, min_box_size)
    n = int(counts.item())
    return boxes[0, :n].clone()


def generate_inference_proposals(cls_scores, anchors, image_size, device,
                                 num_pre_nms=250, score_threshold=0.3,
                                 nms_threshold=0.4, num_post_nms=50,
                                 min_box_size=10):
    """Proposals for inference with NMS (proposal_utils.py:33-59) -> (f32[K,4], f32[K])."""
    boxes, scores, counts = _select(cls_scores, anchors, image_size, num_pre_nms, score_threshold, min_box_size)
    keep, kc = ops.nms_batched(boxes, None, nms_threshold, post_n=int(num_post_nms), counts=counts)
    ob, osc, _ = ops.gather_kept(boxes, scores, keep, kc, want_rois=False)
    n = int(kc.item())
    return ob[0, :n].clone(), osc[0, :n].clone()
