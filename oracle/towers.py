"""TEST INFRASTRUCTURE (parity oracle) - CPU fp32 restatement of the two LightningDOT towers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product path (lightningdot_b200/) never does.

Pinned: oracle/make_golden.py runs the UNMODIFIED reference classes (dvl.models.bi_encoder.BertEncoder /
UniterEncoder imported from /root/reference under oracle/ref_shims.py) on seeded weights and inputs, checks this
restatement against them, and stores the reference outputs under tests/golden/.

Follows:
  text tower   dvl/models/bi_encoder.py:107-123 (BertEncoder.forward) over transformers==2.3.0 BertModel (DVL.yml:180;
               source not under /root/reference - architecture identical to uniter_model/model/layer.py)
  image tower  dvl/models/bi_encoder.py:163-191 (UniterEncoder.forward) -> uniter_model/model/model.py:356-387
  embeddings   uniter_model/model/model.py:233-246 (text), :262-273 + :328-336 (image), :338-354 (concat + gather)
  layer        uniter_model/model/layer.py:75-101 (attention), :111-115, :139-142, :152-156, :166-170
  head         dvl/models/bi_encoder.py:83-88 / :138-143, applied to seq[:, 0, :] (:120-122 / :188-190)
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-12


def gelu(x):
    # uniter_model/model/layer.py:31-37 (erf form, not the tanh approximation)
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def extended_mask(attention_mask, dtype=torch.float32):
    # uniter_model/model/model.py:362-365
    return (1.0 - attention_mask[:, None, None, :].to(dtype)) * -10000.0


def _no_drop(site, x, attention=False):
    return x


def bert_layer(h, ext_mask, sd, p, heads, drop=_no_drop, layer=0):
    """One post-LN transformer layer; p = 'bert.encoder.layer.{i}.'.  drop(site, x): training-mode dropout hook
    (layer.py:93,113,154; identity in eval mode), sites 4 * layer + {0: probabilities, 1: self-output, 2: output}."""
    B, S, H = h.shape
    dh = H // heads

    def split(x):
        return x.view(B, S, heads, dh).permute(0, 2, 1, 3)

    q = split(linear(h, sd, p + "attention.self.query"))
    k = split(linear(h, sd, p + "attention.self.key"))
    v = split(linear(h, sd, p + "attention.self.value"))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh) + ext_mask
    probs = drop(4 * layer, torch.softmax(scores, dim=-1), attention=True)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(B, S, H)
    a = layer_norm(drop(4 * layer + 1, linear(ctx, sd, p + "attention.output.dense")) + h,
                   sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"])
    i = gelu(linear(a, sd, p + "intermediate.dense"))
    return layer_norm(drop(4 * layer + 2, linear(i, sd, p + "output.dense")) + a,
                      sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])


def num_layers(sd):
    n = 0
    while f"bert.encoder.layer.{n}.output.dense.weight" in sd:
        n += 1
    return n


def encoder(h, ext_mask, sd, heads=12, collect=None, drop=_no_drop):
    for i in range(num_layers(sd)):
        h = bert_layer(h, ext_mask, sd, f"bert.encoder.layer.{i}.", heads, drop, i)
        if collect is not None:
            collect.append(h)
    return h


def text_embeddings(sd, input_ids, position_ids):
    e = (F.embedding(input_ids, sd["bert.embeddings.word_embeddings.weight"])
         + F.embedding(position_ids, sd["bert.embeddings.position_embeddings.weight"])
         + sd["bert.embeddings.token_type_embeddings.weight"][0])
    return layer_norm(e, sd["bert.embeddings.LayerNorm.weight"], sd["bert.embeddings.LayerNorm.bias"])


def image_embeddings(sd, img_feat, img_pos_feat):
    p = "bert.img_embeddings."
    im = layer_norm(linear(img_feat, sd, p + "img_linear"), sd[p + "img_layer_norm.weight"], sd[p + "img_layer_norm.bias"])
    ps = layer_norm(linear(img_pos_feat, sd, p + "pos_linear"), sd[p + "pos_layer_norm.weight"], sd[p + "pos_layer_norm.bias"])
    e = im + ps + sd["bert.embeddings.token_type_embeddings.weight"][1]
    return layer_norm(e, sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"])


def projection_head(pooled, sd):
    if "encode_proj.0.weight" not in sd:
        return pooled
    x = gelu(linear(pooled, sd, "encode_proj.0"))
    x = layer_norm(x, sd["encode_proj.2.weight"], sd["encode_proj.2.bias"])
    return linear(x, sd, "encode_proj.3")


SITE_EMB = 0x7E0   # (oracle/dropout.py)


def text_tower(sd, input_ids, attention_mask, position_ids, heads=12, collect=None, drop=_no_drop):
    """-> (sequence_output [B, L, H], pooled [B, D]).  drop: training-mode dropout hook (identity = eval mode); the
    embedding dropout of model.py:245 is site SITE_EMB."""
    h = drop(SITE_EMB, text_embeddings(sd, input_ids, position_ids))
    if collect is not None:
        collect.append(h)
    h = encoder(h, extended_mask(attention_mask), sd, heads, collect, drop)
    return h, projection_head(h[:, 0, :], sd)


def image_tower(sd, input_ids, attention_mask, position_ids, img_feat, img_pos_feat, gather_index=None, heads=12,
                collect=None, drop=_no_drop):
    """-> (sequence_output [B, 1 + R, H], pooled [B, D]).  The reference applies one nn.Dropout to the text embedding and
    one to the region embeddings (model.py:245,272) before concatenating them; with iid masks that is one dropout over
    the concatenated [B, 1 + R, H] input, which is how the hook (site SITE_EMB) is applied."""
    t = text_embeddings(sd, input_ids, position_ids)
    r = image_embeddings(sd, img_feat, img_pos_feat)
    h = torch.cat([t, r], dim=1)
    if gather_index is not None:
        h = torch.gather(h, 1, gather_index.unsqueeze(-1).expand(-1, -1, h.shape[-1]))
    h = drop(SITE_EMB, h)
    if collect is not None:
        collect.append(h)
    h = encoder(h, extended_mask(attention_mask), sd, heads, collect, drop)
    return h, projection_head(h[:, 0, :], sd)
