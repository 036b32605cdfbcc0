"""Device execution of the two towers: sequences of C-ABI kernel launches over 16-bit activations.

One `TowerEngine` per tower holds the inference copies of the weights (16-bit matrices with Q|K|V fused to one
[3H, H] operand, fp32 biases / LayerNorm parameters) and runs

    embeddings -> 12 x [ QKV GEMM -> attention -> O GEMM(+bias+residual) -> LN -> FFN-up GEMM(+bias+GELU)
                         -> FFN-down GEMM(+bias+residual) -> LN ] -> CLS rows -> projection head

exactly as uniter_model/model/model.py:356-387 + layer.py:159-170 + dvl/models/bi_encoder.py:83-88,120-122 do.
All GEMMs are the tcgen05 kernel (ldot_linear); everything runs on torch's current stream.  No torch math is
used on this path.
"""
import os

import torch

from . import _lib

SEQ_BATCH_TOKENS = 524288   # tokens per engine pass (bounds activation scratch to ~10 GB)
# LDOT_FUSED_ATTN=0: Q | K | V projection and attention as two kernels with the [tokens, 3H] round trip (measurement switch)
FUSED_ATTN = os.environ.get("LDOT_FUSED_ATTN", "1") != "0"


def _fmt_of(dtype):
    return {torch.float16: _lib.COARSE_FP16, torch.bfloat16: _lib.COARSE_BF16}[dtype]


class TowerEngine:
    def __init__(self, kind, hidden, heads, ffn, layers, dtype=torch.bfloat16, pre_ln_f32=True, fuse_ln=True):
        assert kind in ("txt", "img")
        if hidden != heads * 64:
            raise ValueError("the attention kernel needs head_dim == 64")
        self.kind, self.H, self.heads, self.ffn, self.layers = kind, hidden, heads, ffn, layers
        self.dtype, self.fmt = dtype, _fmt_of(dtype)
        self.pre_ln_f32 = pre_ln_f32
        # fuse_ln: BertSelfOutput / BertOutput as ONE kernel (GEMM + bias + residual + LayerNorm, cluster of 3 CTAs per
        # row block); otherwise GEMM (+bias +residual, fp32 or 16-bit sums) followed by the LayerNorm kernel
        self.fuse_ln = fuse_ln and hidden % 32 == 0 and hidden <= 768
        self.w = None
        self.signature = None

    # ------------------------------------------------------------------------------------------------ weights
    def load(self, sd, device, trainable=None):
        """Build the device inference copies from a tower state dict (keys of SURVEY.md Appendix B).

        Weights whose fp32 parameter lives in a FusedAdamW flat buffer are ALIASED instead of copied: 16-bit matrices
        become views of the optimiser's 16-bit mirror (written by the AdamW kernel in the same pass as the fp32
        update), fp32 vectors views of the master itself, and Q|K|V one strided view over three adjacent parameters.
        `self.aliased` then tells the owner that optimiser steps need no reload.  `trainable`: names of the parameters
        that can change (None = all)."""
        dt = self.dtype
        device = torch.device(device)
        self.aliased = True

        def note(k, is_alias):
            if not is_alias and (trainable is None or k in trainable):
                self.aliased = False

        def m16_(k, row=None):
            t = sd[k].detach()
            sv = _lib.shadow_view(t, dt) if t.device == device else None
            if sv is not None:
                sv.copy_(t)     # (the mirror may be stale: parameters were loaded / edited since the last step)
                return (sv if row is None else sv[row]), True
            t = t if row is None else t[row]
            return t.to(device=device, dtype=dt).contiguous(), False

        def f32_(k, row=None):
            t = sd[k].detach()
            t = t if row is None else t[row]
            out = t.to(device=device, dtype=torch.float32).contiguous()
            return out, out.data_ptr() == t.data_ptr()

        def m16(k, row=None):
            out, alias = m16_(k, row)
            note(k, alias)
            return out

        def f32(k, row=None):
            out, alias = f32_(k, row)
            note(k, alias)
            return out

        def fused3(keys, conv):
            """[3 n, ...] operand over three parameters: ONE strided view when their (aliased) copies are adjacent."""
            parts, aliases = zip(*[conv(k) for k in keys])
            step = parts[0].numel() * parts[0].element_size()
            same = all(parts[j].untyped_storage().data_ptr() == parts[0].untyped_storage().data_ptr() for j in (1, 2))
            if same and all(aliases) and all(parts[j].data_ptr() == parts[0].data_ptr() + j * step for j in (1, 2)):
                shape = (3 * parts[0].shape[0],) + tuple(parts[0].shape[1:])
                return torch.as_strided(parts[0], shape, parts[0].stride())
            for k in keys:
                note(k, False)
            return torch.cat(parts, 0).contiguous()

        w = {}
        e = "bert.embeddings."
        w["word"], w["pos"] = m16(e + "word_embeddings.weight"), m16(e + "position_embeddings.weight")
        w["type0"] = m16(e + "token_type_embeddings.weight", row=0)
        w["type1_f32"] = f32(e + "token_type_embeddings.weight", row=1)
        w["emb_ln_g"], w["emb_ln_b"] = f32(e + "LayerNorm.weight"), f32(e + "LayerNorm.bias")
        self.vocab, self.max_pos = w["word"].shape[0], w["pos"].shape[0]
        if self.kind == "img":
            p = "bert.img_embeddings."
            w["img_w"], w["img_bias"] = m16(p + "img_linear.weight"), f32(p + "img_linear.bias")
            w["img_ln_g"], w["img_ln_b"] = f32(p + "img_layer_norm.weight"), f32(p + "img_layer_norm.bias")
            w["pos_w"], w["pos_bias"] = f32(p + "pos_linear.weight"), f32(p + "pos_linear.bias")
            w["pos_ln_g"], w["pos_ln_b"] = f32(p + "pos_layer_norm.weight"), f32(p + "pos_layer_norm.bias")
            w["iemb_ln_g"], w["iemb_ln_b"] = f32(p + "LayerNorm.weight"), f32(p + "LayerNorm.bias")
            self.img_dim = w["img_w"].shape[1]
        for i in range(self.layers):
            p = f"bert.encoder.layer.{i}."
            a = p + "attention.self."
            w[f"qkv_w{i}"] = fused3([a + "query.weight", a + "key.weight", a + "value.weight"], m16_)
            w[f"qkv_b{i}"] = fused3([a + "query.bias", a + "key.bias", a + "value.bias"], f32_)
            w[f"o_w{i}"], w[f"o_b{i}"] = m16(p + "attention.output.dense.weight"), f32(p + "attention.output.dense.bias")
            w[f"ln1_g{i}"], w[f"ln1_b{i}"] = f32(p + "attention.output.LayerNorm.weight"), f32(p + "attention.output.LayerNorm.bias")
            w[f"f1_w{i}"], w[f"f1_b{i}"] = m16(p + "intermediate.dense.weight"), f32(p + "intermediate.dense.bias")
            w[f"f2_w{i}"], w[f"f2_b{i}"] = m16(p + "output.dense.weight"), f32(p + "output.dense.bias")
            w[f"ln2_g{i}"], w[f"ln2_b{i}"] = f32(p + "output.LayerNorm.weight"), f32(p + "output.LayerNorm.bias")
        self.project = "encode_proj.0.weight" in sd
        if self.project:
            w["p0_w"], w["p0_b"] = m16("encode_proj.0.weight"), f32("encode_proj.0.bias")
            w["p_ln_g"], w["p_ln_b"] = f32("encode_proj.2.weight"), f32("encode_proj.2.bias")
            w["p3_w"], w["p3_b"] = m16("encode_proj.3.weight"), f32("encode_proj.3.bias")
            self.out_dim = w["p3_w"].shape[0]
        else:
            self.out_dim = self.H
        self.w = w
        self.device = device

    # ------------------------------------------------------------------------------------------------ kernels
    def _linear(self, a, lda, wt, bias, out, M, act=0, residual=None, rows_k=None):
        lib = _lib.load()
        N, K = wt.shape
        _lib.check(lib.ldot_linear(_lib.ptr(a), lda, _lib.ptr(wt), K, _lib.ptr(bias), _lib.ptr(residual),
                                   0 if residual is None else residual.stride(0), _lib.ptr(out), out.stride(0),
                                   M, N, K, self.fmt, act, int(out.dtype == torch.float32), _lib.stream_ptr()))

    def _linear_ln(self, a, lda, wt, bias, residual, gamma, beta, out, M):
        lib = _lib.load()
        N, K = wt.shape
        _lib.check(lib.ldot_linear_ln(_lib.ptr(a), lda, _lib.ptr(wt), K, _lib.ptr(bias), _lib.ptr(residual),
                                      residual.stride(0), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(out), out.stride(0),
                                      M, N, K, self.fmt, _lib.stream_ptr()))

    def _layernorm(self, x, g, b, out, rows, H):
        lib = _lib.load()
        _lib.check(lib.ldot_layernorm(_lib.ptr(x), x.stride(0), int(x.dtype == torch.float32), _lib.ptr(g), _lib.ptr(b),
                                      _lib.ptr(out), out.stride(0), rows, H, self.fmt, _lib.stream_ptr()))

    def _layers(self, h, mask, B, S, cls_only=False):
        """h: [B*S, H] 16-bit (updated in place); mask int64 [B, S].  Returns (h, cls) where cls is None, or - with
        cls_only - the compact [B, H] matrix of last-layer [CLS] rows.  With cls_only the last layer is evaluated for the
        [CLS] query position alone (keys / values still come from every position): the rows the reference computes and
        then discards at bi_encoder.py:120,188 are not produced, the [CLS] row is bit-identical."""
        lib = _lib.load()
        T, H, dt, dev = B * S, self.H, self.dtype, h.device
        w = self.w
        qkv = None   # (only the unfused layers need the [T, 3H] projection buffer)
        ctx = torch.empty((T, H), dtype=dt, device=dev)
        pre = None if self.fuse_ln else torch.empty((T, H), dtype=torch.float32 if self.pre_ln_f32 else dt, device=dev)
        a = torch.empty((T, H), dtype=dt, device=dev)
        f = torch.empty((T, self.ffn), dtype=dt, device=dev)
        stream = _lib.stream_ptr()

        def block(i, x_in, ldx, ctx_, a_, f_, out_, rows, q_rows, attend=True):
            """attention output ctx_ -> BertSelfOutput -> BertIntermediate -> BertOutput for `rows` rows; x_in (row pitch
            ldx) is the layer input of those rows (the residual)."""
            if attend:
                _lib.check(lib.ldot_attention(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx_), B, S, H, self.heads, q_rows,
                                              self.fmt, stream))
            res = x_in.as_strided((rows, H), (ldx, 1))
            if self.fuse_ln:
                self._linear_ln(ctx_, H, w[f"o_w{i}"], w[f"o_b{i}"], res, w[f"ln1_g{i}"], w[f"ln1_b{i}"], a_, rows)
                self._linear(a_, H, w[f"f1_w{i}"], w[f"f1_b{i}"], f_, rows, act=1)
                self._linear_ln(f_, self.ffn, w[f"f2_w{i}"], w[f"f2_b{i}"], a_, w[f"ln2_g{i}"], w[f"ln2_b{i}"], out_, rows)
                return
            pre_ = pre[:rows]
            self._linear(ctx_, H, w[f"o_w{i}"], w[f"o_b{i}"], pre_, rows, residual=res)
            self._layernorm(pre_, w[f"ln1_g{i}"], w[f"ln1_b{i}"], a_, rows, H)
            self._linear(a_, H, w[f"f1_w{i}"], w[f"f1_b{i}"], f_, rows, act=1)
            self._linear(f_, self.ffn, w[f"f2_w{i}"], w[f"f2_b{i}"], pre_, rows, residual=a_)
            self._layernorm(pre_, w[f"ln2_g{i}"], w[f"ln2_b{i}"], out_, rows, H)

        for i in range(self.layers):
            last_cls = cls_only and i == self.layers - 1
            if FUSED_ATTN and not last_cls:
                # BertSelfAttention as ONE kernel: the projection's [T, 3H] output stays in shared memory / TMEM
                _lib.check(lib.ldot_qkv_attention(_lib.ptr(h), H, _lib.ptr(w[f"qkv_w{i}"]), H, _lib.ptr(w[f"qkv_b{i}"]),
                                                  _lib.ptr(mask), _lib.ptr(ctx), B, S, H, self.heads, H, self.fmt, stream))
                block(i, h, H, ctx, a, f, h, T, S, attend=False)
                continue
            if qkv is None:
                qkv = torch.empty((T, 3 * H), dtype=dt, device=dev)
            self._linear(h, H, w[f"qkv_w{i}"], w[f"qkv_b{i}"], qkv, T)
            if last_cls:
                cls = torch.empty((B, H), dtype=dt, device=dev)
                block(i, h, S * H, ctx[:B], a[:B], f[:B], cls, B, 1)
                return h, cls
            block(i, h, H, ctx, a, f, h, T, S)
        return h, None

    def _head(self, h, B, S):
        """CLS rows (row pitch S*H; S = 1 for the compact matrix) -> projection head -> fp32 [B, out_dim]."""
        w, H, dev = self.w, self.H, h.device
        if not self.project:
            return h.view(B, S, H)[:, 0, :].float()
        x = torch.empty((B, 2 * H), dtype=self.dtype, device=dev)
        self._linear(h, S * H, w["p0_w"], w["p0_b"], x, B, act=1)
        y = torch.empty((B, 2 * H), dtype=self.dtype, device=dev)
        self._layernorm(x, w["p_ln_g"], w["p_ln_b"], y, B, 2 * H)
        out = torch.empty((B, self.out_dim), dtype=torch.float32, device=dev)
        self._linear(y, 2 * H, w["p3_w"], w["p3_b"], out, B)
        return out

    # ------------------------------------------------------------------------------------------------ towers
    def _embed_text(self, ids, pos_ids, out, B, L, out_seq):
        lib = _lib.load()
        w = self.w
        stride = 0 if pos_ids.shape[0] == 1 else pos_ids.stride(0)
        _lib.check(lib.ldot_embed_text(_lib.ptr(ids), _lib.ptr(pos_ids), stride, _lib.ptr(w["word"]), _lib.ptr(w["pos"]),
                                       _lib.ptr(w["type0"]), _lib.ptr(w["emb_ln_g"]), _lib.ptr(w["emb_ln_b"]),
                                       _lib.ptr(out), B, L, out_seq, self.H, self.vocab, self.max_pos, self.fmt,
                                       _lib.stream_ptr()))

    @staticmethod
    def _i64(t, dev):
        return t.to(device=dev, dtype=torch.int64).contiguous()

    def encode_text(self, input_ids, attention_mask, position_ids, want_seq=False):
        """-> (sequence_output [B, L, H] 16-bit or None, pooled [B, D] fp32)"""
        dev = self.device
        ids, mask, pos = self._i64(input_ids, dev), self._i64(attention_mask, dev), self._i64(position_ids, dev)
        B, L = ids.shape
        if L > 128:
            raise ValueError(f"sequence length {L} > 128 is not supported by the attention kernel")
        if pos.dim() == 1:
            pos = pos[None, :]
        step = max(1, SEQ_BATCH_TOKENS // L)
        pooled, seqs = [], []
        for b0 in range(0, B, step):
            b1 = min(B, b0 + step)
            nb = b1 - b0
            h = torch.empty((nb * L, self.H), dtype=self.dtype, device=dev)
            self._embed_text(ids[b0:b1], pos if pos.shape[0] == 1 else pos[b0:b1], h, nb, L, L)
            h, cls = self._layers(h, mask[b0:b1], nb, L, cls_only=not want_seq)
            pooled.append(self._head(h, nb, L) if cls is None else self._head(cls, nb, 1))
            if want_seq:
                seqs.append(h.view(nb, L, self.H))
        return (torch.cat(seqs, 0) if want_seq else None), (pooled[0] if len(pooled) == 1 else torch.cat(pooled, 0))

    def encode_image(self, input_ids, attention_mask, position_ids, img_feat, img_pos_feat, gather_index=None,
                     want_seq=False):
        """-> (sequence_output [B, Lt + R, H] 16-bit or None, pooled [B, D] fp32)"""
        lib = _lib.load()
        dev, w, H = self.device, self.w, self.H
        ids, mask, pos = self._i64(input_ids, dev), self._i64(attention_mask, dev), self._i64(position_ids, dev)
        feat = img_feat.to(device=dev, dtype=torch.float32).contiguous()
        box = img_pos_feat.to(device=dev, dtype=torch.float32).contiguous()
        B, Lt = ids.shape
        R = feat.shape[1]
        S = Lt + R
        if S > 128:
            raise ValueError(f"sequence length {S} > 128 is not supported by the attention kernel")
        if mask.shape[1] != S:
            raise ValueError(f"attention_mask has {mask.shape[1]} positions, expected {S}")
        if gather_index is not None:
            gi = gather_index.to(dev)
            if not torch.equal(gi, torch.arange(S, device=dev)[None, :].expand(B, S)):
                raise NotImplementedError("only the identity gather_index of dvl/data/itm.py is supported")
        if pos.dim() == 1:
            pos = pos[None, :]
        step = max(1, SEQ_BATCH_TOKENS // S)
        pooled, seqs = [], []
        stream = _lib.stream_ptr()
        for b0 in range(0, B, step):
            b1 = min(B, b0 + step)
            nb = b1 - b0
            h = torch.empty((nb * S, H), dtype=self.dtype, device=dev)
            self._embed_text(ids[b0:b1], pos if pos.shape[0] == 1 else pos[b0:b1], h, nb, Lt, S)
            f16 = torch.empty((nb * R, self.img_dim), dtype=self.dtype, device=dev)
            _lib.check(lib.ldot_cast_f32(_lib.ptr(feat[b0:b1]), _lib.ptr(f16), nb * R * self.img_dim, self.fmt, stream))
            lin = torch.empty((nb * R, H), dtype=torch.float32, device=dev)
            self._linear(f16, self.img_dim, w["img_w"], w["img_bias"], lin, nb * R)
            _lib.check(lib.ldot_embed_image(
                _lib.ptr(lin), _lib.ptr(box[b0:b1]), _lib.ptr(w["img_ln_g"]), _lib.ptr(w["img_ln_b"]), _lib.ptr(w["pos_w"]),
                _lib.ptr(w["pos_bias"]), _lib.ptr(w["pos_ln_g"]), _lib.ptr(w["pos_ln_b"]), _lib.ptr(w["type1_f32"]),
                _lib.ptr(w["iemb_ln_g"]), _lib.ptr(w["iemb_ln_b"]), _lib.ptr(h), nb, R, S, Lt, H, self.fmt, stream))
            h, cls = self._layers(h, mask[b0:b1], nb, S, cls_only=not want_seq)
            pooled.append(self._head(h, nb, S) if cls is None else self._head(cls, nb, 1))
            if want_seq:
                seqs.append(h.view(nb, S, H))
        return (torch.cat(seqs, 0) if want_seq else None), (pooled[0] if len(pooled) == 1 else torch.cat(pooled, 0))
