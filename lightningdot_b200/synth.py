"""Seeded synthetic inputs for tests and benchmarks.

There is no network for checkpoints or LMDB features, so weights are random-init with the reference's initialiser
(N(0, 0.02) for Linear / Embedding weights, LayerNorm gamma = 1, beta = 0, Linear bias = 0 -
uniter_model/model/model.py:134-147) and inputs are drawn to the shapes of dvl/data/itm.py:203-288.
Everything is generated on the CPU generator with a fixed seed so that every machine sees identical tensors.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

VOCAB = 28996          # config/img_base.json:13 (bert-base-cased)
HIDDEN = 768
FFN = 3072
HEADS = 12
LAYERS = 12
IMG_DIM = 2048         # dvl/const.py:1
MAX_POS = 512
TYPE_VOCAB = 2


def tower_param_shapes(kind, hidden=HIDDEN, ffn=FFN, layers=LAYERS, vocab=VOCAB, project_dim=768, img_dim=IMG_DIM):
    """Ordered {state-dict key: shape} of one tower.  kind: 'txt' (BertEncoder) or 'img'
    (UniterEncoder)."""
    s = OrderedDict()
    s["bert.embeddings.word_embeddings.weight"] = (vocab, hidden)
    s["bert.embeddings.position_embeddings.weight"] = (MAX_POS, hidden)
    s["bert.embeddings.token_type_embeddings.weight"] = (TYPE_VOCAB, hidden)
    s["bert.embeddings.LayerNorm.weight"] = (hidden,)
    s["bert.embeddings.LayerNorm.bias"] = (hidden,)
    if kind == "img":
        p = "bert.img_embeddings."
        s[p + "img_linear.weight"] = (hidden, img_dim)
        s[p + "img_linear.bias"] = (hidden,)
        s[p + "img_layer_norm.weight"] = (hidden,)
        s[p + "img_layer_norm.bias"] = (hidden,)
        s[p + "pos_layer_norm.weight"] = (hidden,)
        s[p + "pos_layer_norm.bias"] = (hidden,)
        s[p + "pos_linear.weight"] = (hidden, 7)
        s[p + "pos_linear.bias"] = (hidden,)
        s[p + "mask_embedding.weight"] = (2, img_dim)
        s[p + "LayerNorm.weight"] = (hidden,)
        s[p + "LayerNorm.bias"] = (hidden,)
    for i in range(layers):
        p = f"bert.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            s[p + f"attention.self.{nm}.weight"] = (hidden, hidden)
            s[p + f"attention.self.{nm}.bias"] = (hidden,)
        s[p + "attention.output.dense.weight"] = (hidden, hidden)
        s[p + "attention.output.dense.bias"] = (hidden,)
        s[p + "attention.output.LayerNorm.weight"] = (hidden,)
        s[p + "attention.output.LayerNorm.bias"] = (hidden,)
        s[p + "intermediate.dense.weight"] = (ffn, hidden)
        s[p + "intermediate.dense.bias"] = (ffn,)
        s[p + "output.dense.weight"] = (hidden, ffn)
        s[p + "output.dense.bias"] = (hidden,)
        s[p + "output.LayerNorm.weight"] = (hidden,)
        s[p + "output.LayerNorm.bias"] = (hidden,)
    s["bert.pooler.dense.weight"] = (hidden, hidden)
    s["bert.pooler.dense.bias"] = (hidden,)
    if project_dim > 0:
        s["encode_proj.0.weight"] = (2 * hidden, hidden)
        s["encode_proj.0.bias"] = (2 * hidden,)
        s["encode_proj.2.weight"] = (2 * hidden,)
        s["encode_proj.2.bias"] = (2 * hidden,)
        s["encode_proj.3.weight"] = (project_dim, 2 * hidden)
        s["encode_proj.3.bias"] = (project_dim,)
    return s


def random_tower_state(kind, seed=42, perturb=False, **shape_kw):
    """Random state dict of one tower.  perturb=False is the reference initialiser; perturb=True additionally
    draws non-trivial biases and LayerNorm affine parameters (a 'trained-like' state) so that parity tests
    exercise every term of the arithmetic."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, shp in tower_param_shapes(kind, **shape_kw).items():
        is_ln = "LayerNorm" in k or "layer_norm" in k or k.startswith("encode_proj.2.")
        if is_ln and k.endswith("weight"):
            t = torch.ones(shp)
            if perturb:
                t = t + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = torch.zeros(shp)
            if perturb:
                t = 0.05 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        out[k] = t
    if "bert.embeddings.word_embeddings.weight" in out and not perturb:
        pass  # HF zeroes the padding_idx row at init; row 0 is never looked up un-masked, keep N(0, 0.02)
    return out


def conditioned_tower_state(kind, seed=42, head_scale=0.1, **shape_kw):
    """random_tower_state(perturb=True) with the last projection (encode_proj.3) scaled by `head_scale`.  Random-init
    embeddings have norm ~22, so in-batch scores differ by several units and the softmax of the loss turns 16-bit score
    noise (0.2 units in bf16) into ~15 % gradient changes - a property of the random model, not of the kernels.  Scaling
    the head by 0.1 divides the scores by 100: the loss keeps every term of its arithmetic but is well conditioned, so a
    whole training step can be compared with the reference at the precision of the backward kernels themselves."""
    sd = random_tower_state(kind, seed=seed, perturb=True, **shape_kw)
    sd["encode_proj.3.weight"] = sd["encode_proj.3.weight"] * head_scale
    sd["encode_proj.3.bias"] = sd["encode_proj.3.bias"] * head_scale
    return sd


def text_batch(batch, seq_len=32, seed=0, ragged=False, min_len=8, vocab=VOCAB):
    """'txts' sub-batch of itm_fast_collate (dvl/data/itm.py:230-246): [CLS]=101 ... [SEP]=102, zero padded."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, vocab, (batch, seq_len), generator=g)
    if ragged:
        lens = torch.randint(min_len, seq_len + 1, (batch,), generator=g)
    else:
        lens = torch.full((batch,), seq_len, dtype=torch.long)
    ar = torch.arange(seq_len)[None, :]
    mask = (ar < lens[:, None]).long()
    ids = ids * mask
    ids[:, 0] = 101
    ids[torch.arange(batch), lens - 1] = 102
    return {
        "input_ids": ids,
        "position_ids": torch.arange(seq_len, dtype=torch.long)[None, :],
        "attention_mask": mask,
        "img_feat": None, "img_pos_feat": None, "img_masks": None, "gather_index": None,
    }


def image_batch(batch, num_bb=36, seed=0, ragged=False, min_bb=10, img_dim=IMG_DIM):
    """'imgs' sub-batch of itm_fast_collate (dvl/data/itm.py:248-262): one [CLS] token + R region features."""
    g = torch.Generator().manual_seed(seed + 7919)
    feat = torch.relu(torch.randn(batch, num_bb, img_dim, generator=g)) * 0.5
    xy = torch.rand(batch, num_bb, 2, generator=g) * 0.7
    wh = torch.rand(batch, num_bb, 2, generator=g) * 0.25 + 0.05
    pos = torch.cat([xy, xy + wh, wh, wh[..., :1] * wh[..., 1:]], dim=-1)
    if ragged:
        nbb = torch.randint(min_bb, num_bb + 1, (batch,), generator=g)
    else:
        nbb = torch.full((batch,), num_bb, dtype=torch.long)
    ar = torch.arange(num_bb)[None, :]
    valid = (ar < nbb[:, None])
    feat = feat * valid[..., None]
    pos = pos * valid[..., None]
    mask = torch.cat([torch.ones(batch, 1, dtype=torch.long), valid.long()], dim=1)
    return {
        "input_ids": torch.full((batch, 1), 101, dtype=torch.long),
        "position_ids": torch.zeros(1, 1, dtype=torch.long),
        "attention_mask": mask,
        "img_feat": feat, "img_pos_feat": pos, "img_masks": None,
        "gather_index": torch.arange(1 + num_bb, dtype=torch.long)[None, :].repeat(batch, 1),
    }


def gaussian_index(n, d=768, seed=42):
    """Index-boundary fixture: X ~ N(0, 1) / sqrt(d), fp32 [n, d]."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, d), dtype=np.float32) / np.float32(math.sqrt(d))).astype(np.float32)


def collinear_index(n, d=768, seed=42, spread=0.23):
    """Stress fixture mimicking random-init towers: x = c + spread * eps (pairwise cosine ~ 0.95)."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((1, d), dtype=np.float32) / np.float32(math.sqrt(d))
    e = rng.standard_normal((n, d), dtype=np.float32) / np.float32(math.sqrt(d))
    return (c + np.float32(spread) * e).astype(np.float32)


def planted_queries(x, nq, sigma=1.0, seed=43, scale=1.0):
    """Queries q_j = scale * (x[gt_j] + sigma * N(0,1)/sqrt(d)); returns (q, gt).  sigma tunes how hard the
    retrieval is (Recall@1 well inside (0, 1) so that ranking errors would show)."""
    rng = np.random.default_rng(seed)
    n, d = x.shape
    gt = rng.integers(0, n, size=nq)
    noise = rng.standard_normal((nq, d), dtype=np.float32) / np.float32(math.sqrt(d))
    q = (x[gt] + np.float32(sigma) * noise) * np.float32(scale)
    return q.astype(np.float32), gt.astype(np.int64)


# ---------------------------------------------------------------------------------------------------------------
# The itm_fast_collate batch (dvl/data/itm.py:203-288) - the input contract of BiEncoder.forward
# ---------------------------------------------------------------------------------------------------------------
def itm_samples(batch, seq_len=32, num_bb=36, seed=0, ragged=True):
    """Per-sample tuples exactly as the reference's ItmFastDataset.__getitem__ hands them to itm_fast_collate
    (dvl/data/itm.py:68-131): (input_ids [len], img_feat [nbb, 2048], img_pos_feat [nbb, 7], img_input_ids [1],
    attn_masks_text [len], attn_masks_img [1 + nbb], txt id, image file name, neg_imgs, neg_txts, caption_ids,
    attn_masks_captions) - un-padded, no hard negatives, no captions.  Caption j belongs to image j // 5."""
    tb = text_batch(batch, seq_len, seed=seed, ragged=ragged)
    ib = image_batch(batch, num_bb, seed=seed, ragged=ragged)
    out = []
    for j in range(batch):
        tl = int(tb["attention_mask"][j].sum())
        nbb = int(ib["attention_mask"][j].sum()) - 1
        out.append((tb["input_ids"][j, :tl].clone(), ib["img_feat"][j, :nbb].clone(), ib["img_pos_feat"][j, :nbb].clone(),
                    torch.tensor([101], dtype=torch.long), torch.ones(tl, dtype=torch.long),
                    torch.ones(nbb + 1, dtype=torch.long), str(j), f"img_{j // 5:07d}.npz", None, None, None, None))
    return out


def itm_batch(batch, seq_len=32, num_bb=36, seed=0, ragged=True):
    """The nested batch itm_fast_collate builds from itm_samples(...) (same arguments): tensors padded to the longest
    member of the batch, `position_ids` [1, L], identity `gather_index`, the bookkeeping lists.  Pinned against the
    reference's own collate function by oracle/make_golden.py (tests/golden/itm_batch_schema.json)."""
    tb = text_batch(batch, seq_len, seed=seed, ragged=ragged)
    ib = image_batch(batch, num_bb, seed=seed, ragged=ragged)
    tl = int(tb["attention_mask"].sum(1).max())
    nbb = int(ib["attention_mask"].sum(1).max()) - 1
    none5 = {"img_feat": None, "img_pos_feat": None, "img_masks": None, "gather_index": None}
    txts = {"input_ids": tb["input_ids"][:, :tl].contiguous(), "position_ids": torch.arange(tl, dtype=torch.long)[None, :],
            "attention_mask": tb["attention_mask"][:, :tl].contiguous(), **none5}
    imgs = {"input_ids": ib["input_ids"], "position_ids": torch.zeros(1, 1, dtype=torch.long),
            "attention_mask": ib["attention_mask"][:, :nbb + 1].contiguous(),
            "img_feat": ib["img_feat"][:, :nbb].contiguous(), "img_pos_feat": ib["img_pos_feat"][:, :nbb].contiguous(),
            "img_masks": None, "gather_index": torch.arange(nbb + 1, dtype=torch.long)[None, :].repeat(batch, 1)}
    caps = {"input_ids": None, "position_ids": None, "attention_mask": None, **none5}
    return {"txts": txts, "imgs": imgs, "caps": caps, "sample_size": batch, "pos_ctx_indices": list(range(batch)),
            "neg_ctx_indices": [], "txt_index": [str(j) for j in range(batch)],
            "img_fname": [f"img_{j // 5:07d}.npz" for j in range(batch)]}


def describe_batch(b):
    """Structure + content fingerprint of a nested batch: {path: [dtype, shape, sha1 of the bytes] | value}."""
    import hashlib
    out = {}

    def walk(prefix, v):
        if isinstance(v, dict):
            for k, x in v.items():
                walk(f"{prefix}.{k}" if prefix else k, x)
        elif torch.is_tensor(v):
            out[prefix] = [str(v.dtype), list(v.shape), hashlib.sha1(v.contiguous().numpy().tobytes()).hexdigest()]
        else:
            out[prefix] = v
    walk("", b)
    return out


# ---------------------------------------------------------------------------------------------------------------
# Synthetic ITM databases on disk (the reference's text / image database layout, lightningdot_b200/data.py)
# ---------------------------------------------------------------------------------------------------------------
def make_itm_db(root, n_img, caps_per_img=5, seq_len=32, num_bb=36, seed=0, txt2img=None, name="val", conf_th=0.2,
                max_bb=100, min_bb=10, compress=False, write_images=True):
    """Write <root>/txt_<name>.db (tokenised captions) and <root>/img/ (region features) for n_img images and
    n_img * caps_per_img captions drawn by text_batch / image_batch(ragged=True).  Caption j describes image
    j // caps_per_img unless `txt2img` (list of image numbers, one per caption) says otherwise.  Image features are
    stored as fp16 (like the reference's converter); boxes past an image's own count carry confidence 0 so that the
    confidence-threshold rule (conf_th, min_bb, max_bb) recovers exactly that count.  -> (txt_db dir, img_db dir)."""
    import os
    from . import data
    n_cap = n_img * caps_per_img
    tb = text_batch(n_cap, seq_len, seed=seed, ragged=True)
    lens = tb["attention_mask"].sum(1).tolist()
    img_name = [f"img_{i:07d}.npz" for i in range(n_img)]
    owner = [j // caps_per_img for j in range(n_cap)] if txt2img is None else [int(v) for v in txt2img]
    records = {}
    for j in range(n_cap):
        body = tb["input_ids"][j, 1:lens[j] - 1].tolist()     # [CLS] / [SEP] are re-added by TxtTokLmdb.combine_inputs
        records[str(j)] = {"input_ids": body, "img_fname": img_name[owner[j]], "id": str(j)}
    txt_dir = data.write_txt_db(os.path.join(root, f"txt_{name}.db"), records, backend="flat")
    img_dir = os.path.join(root, "img")
    if write_images:
        ib = image_batch(n_img, num_bb, seed=seed, ragged=True, min_bb=min_bb)
        nbb = (ib["attention_mask"].sum(1) - 1).tolist()
        feats = {}
        for i in range(n_img):
            conf = np.zeros(num_bb, dtype=np.float32)
            conf[:nbb[i]] = 0.9
            feats[img_name[i]] = {"features": ib["img_feat"][i].numpy(), "norm_bb": ib["img_pos_feat"][i, :, :6].numpy(),
                                  "conf": conf}
        data.write_img_db(img_dir, feats, conf_th=conf_th, max_bb=max_bb, min_bb=min_bb, num_bb=num_bb,
                          compress=compress, backend="flat")
    return txt_dir, img_dir
