"""TEST INFRASTRUCTURE - planted-margin labels for the end-to-end Recall@1/5/10 equality fixtures (SURVEY.md section 7,
hard part 2 (iii)).

Random-init towers give nearly collinear embeddings (pairwise cosine 0.98), so most rank boundaries are separated by
less than the rounding noise of 16-bit towers and "identical Recall@k" is not a property any 16-bit implementation can
have on arbitrary labels.  Recall only depends on where each caption's LABELLED image lands relative to the 1|2, 5|6 and
10|11 rank boundaries (dvl/trainer.py:172-188), and on a synthetic database the labels are free: given the reference's
fp32 score matrix S, every caption is assigned an image whose position is separated from those boundaries by at least
`delta` in BOTH directions of retrieval.  Any implementation whose score errors stay below delta / 2 must then reproduce
Recall@1/5/10 exactly; an implementation with wrong towers or a wrong search does not.

  plan(S, delta)        -> owner image per caption, class per caption (0: hit@1, 1: hit@5 only, 2: hit@10 only, 3: miss)
  recalls(S, owner)     -> (recall_txt, recall_img) as eval_model_on_dataloader computes them, from a score matrix
"""
import numpy as np


def plan(S, delta, want=(0.3, 0.2, 0.2, 0.3), seed=0):
    n_cap, n_img = S.shape
    order = np.argsort(-S, 1, kind="stable")
    srt = np.take_along_axis(S, order, 1)
    csrt = -np.sort(-S, 0)
    # image -> text direction: pair (j, i) is safe when S[j, i] is delta away from column i's 1|2, 5|6, 10|11 boundaries
    margin = np.full(S.shape, np.inf, dtype=np.float32)
    for t in (1, 5, 10):
        hi, lo = csrt[t - 1][None], csrt[t][None]
        margin = np.minimum(margin, np.where(S >= hi, S - lo, hi - S))
    rng = np.random.default_rng(seed)
    target = rng.choice(4, size=n_cap, p=want)
    owner = np.full(n_cap, -1, dtype=np.int64)
    cls = np.full(n_cap, -1, dtype=np.int64)
    counts = np.zeros(n_img, dtype=np.int64)
    forced = 0
    for j in range(n_cap):
        s = srt[j]
        cands = {0: [0] if s[0] - s[1] >= delta else [],
                 1: [r for r in range(1, 5) if s[0] - s[r] >= delta and s[r] - s[5] >= delta],
                 2: [r for r in range(5, 10) if s[4] - s[r] >= delta and s[r] - s[10] >= delta]}
        misses = [r for r in range(10, n_img) if s[9] - s[r] >= delta]
        misses.sort(key=lambda r: counts[order[j, r]])        # spread the misses over rarely used images
        pick = None
        for c in [target[j]] + [c for c in (0, 1, 2) if c != target[j]] + [3]:
            rs = misses if c == 3 else cands[c]
            hit = next((r for r in rs if margin[j, order[j, r]] >= delta), None)
            if hit is not None:
                pick = (hit, c)
                break
        if pick is None:   # a "hub" caption that sits near the top of every image's list: best effort
            hit = max(misses, key=lambda r: margin[j, order[j, r]])
            pick, forced = (hit, 3), forced + 1
        owner[j], cls[j] = order[j, pick[0]], pick[1]
        counts[owner[j]] += 1
    return owner, cls, forced


def recalls(S, owner):
    order = np.argsort(-S, 1, kind="stable")
    rank = (order == owner[:, None]).argmax(1)
    recall_txt = {t: float((rank < t).mean()) for t in (1, 5, 10)}
    corder = np.argsort(-S, 0, kind="stable")
    imgs = np.unique(owner)
    recall_img = {t: float(np.mean([np.isin(corder[:t, i], np.nonzero(owner == i)[0]).any() for i in imgs]))
                  for t in (1, 5, 10)}
    return recall_txt, recall_img
