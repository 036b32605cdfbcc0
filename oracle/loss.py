"""TEST INFRASTRUCTURE (parity oracle) - in-batch-negative NLL of the bi-encoder, CPU fp32.

Pinned by oracle/make_golden.py against the reference's own BiEncoderNllLoss / _calc_loss (imported from
/root/reference) -> tests/golden/loss_*.npz.

Follows dvl/models/bi_encoder.py:54-68 (dot_product_scores), :615-656 (BiEncoderNllLoss.calc) and
dvl/utils.py:158-167 (_calc_loss, world size 1 branch); symmetric use in train_itm.py:203-222.
"""
import torch
import torch.nn.functional as F


def dot_product_scores(q, ctx):
    return torch.matmul(q, ctx.transpose(0, 1))


def nll(q, ctx, positive_idx, caption_vectors=None, caption_score_weight=0.1, reduction="mean"):
    """-> (loss, correct_count, scores) exactly as BiEncoderNllLoss.calc returns them."""
    scores = dot_product_scores(q, ctx)
    if caption_vectors is not None and caption_score_weight != 0:
        scores = (1 - caption_score_weight) * scores + caption_score_weight * dot_product_scores(q, caption_vectors)
    scores = scores.view(q.size(0), -1)
    logp = F.log_softmax(scores, dim=1)
    target = torch.as_tensor(positive_idx, dtype=torch.long)
    loss = F.nll_loss(logp, target, reduction=reduction)
    correct = (logp.argmax(dim=1) == target).sum()
    return loss, correct, scores


def symmetric_nll(txt, img):
    """train_itm.py:203-222: 0.5 * nll(img -> txt) + 0.5 * nll(txt -> img), positives on the diagonal."""
    pos = list(range(txt.size(0)))
    l_txt, c_txt, _ = nll(img, txt, pos)
    l_img, c_img, _ = nll(txt, img, pos)
    return 0.5 * l_txt + 0.5 * l_img, (c_txt + c_img) / 2
