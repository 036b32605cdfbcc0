"""Training step of the two towers on the device: forward with saved activations, hand-written backward, and the
torch.autograd.Function that lets train_itm.py's `loss.backward()` drive it.

The reference gets its gradients from torch autograd over apex-amp fp16 modules (train_itm.py:252-267,
dvl/models/bi_encoder.py:589-601).  Here one autograd node per tower call owns the whole tower:

    forward    embeddings -> 12 x [QKV GEMM, attention, O GEMM + residual, LayerNorm, FFN-up GEMM, GELU, FFN-down GEMM
               + residual, LayerNorm] -> [CLS] rows -> projection head, keeping 16-bit activations per layer
               (x, qkv, ctx, pre-LN sums, FFN pre-activation and activation: 24.6 KB per token per layer)
    backward   the mirror image: LayerNorm backward (+ bias-gradient column sums) -> wgrad (split-K tcgen05 GEMM that
               reads dY and X as stored, accumulating fp32) and dgrad (tcgen05 GEMM that reads the nn.Linear weight as
               stored; residual-gradient add and GELU' fused in its epilogue) -> attention backward -> ... ->
               embedding backward (scatter-add)

Activations and activation gradients are bf16 (fp16 when setup_for_distributed_mode(fp16=True) selected it - without
loss scaling fp16 gradients underflow, so bf16 is the default and the recommended training dtype); parameter gradients
are fp32 under the reference's parameter names, so torch optimisers, clip_grad_norm_ and state_dict() keep working.
Dropout (hidden_dropout_prob / attention_probs_dropout_prob, bi_encoder.py:97-99) is applied when the module is in
train() mode: after the embedding LayerNorm, on the attention probabilities, and on the BertSelfOutput / BertOutput dense
outputs before the residual add (uniter_model/model/layer.py:93,113,154; model.py:245,272).  Masks are a counter-based
function of (seed drawn from torch's CPU generator per tower call, site, element index) - csrc/dropout.cuh - so the
backward kernels regenerate them instead of storing them.  eval() mode (or probabilities 0) runs the deterministic network.
"""
import weakref

import torch

from . import _lib


class _Tape(object):
    """Activations one forward keeps for its backward."""
    __slots__ = ("kind", "B", "S", "Lt", "R", "ids", "pos", "mask", "layers", "h_last", "x0", "x1", "x2", "feat16",
                 "lin", "box", "h0", "p_hidden", "p_attn", "seed")


SITE_EMB = 0x7E0   # dropout site of the embedding output; layer l uses 4 l + {0: probabilities, 1: self-output, 2: output}


class _GradSinks(dict):
    """{parameter name: fp32 gradient} being produced by one backward.  z(name, shape) hands out the tensor the kernels
    accumulate into: the parameter's existing .grad when the caller supplied it as a sink, else fresh zeros."""

    def __init__(self, sinks, device):
        super().__init__()
        self.sinks, self.device = sinks, device
        self._pending = []

    def z(self, name, *shape):
        t = self.sinks.get(name)
        if t is None:
            t = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self[name] = t
        return t

    def z3(self, names, rows, *rest):
        """One [3 rows, ...] accumulator for the fused Q|K|V gradient: a strided view over the three sinks when they are
        adjacent in memory (FusedAdamW's flat layout), else a temporary that scatter3() adds / hands out afterwards."""
        ts = [self.sinks.get(n) for n in names]
        if all(t is not None for t in ts):
            step = ts[0].numel() * 4
            same = all(ts[j].untyped_storage().data_ptr() == ts[0].untyped_storage().data_ptr() for j in (1, 2))
            if same and all(ts[j].data_ptr() == ts[0].data_ptr() + j * step for j in (1, 2)):
                for n, t in zip(names, ts):
                    self[n] = t
                return torch.as_strided(ts[0], (3 * rows,) + tuple(rest), ts[0].stride())
        tmp = torch.zeros((3 * rows,) + tuple(rest), dtype=torch.float32, device=self.device)
        self._pending.append((names, rows, tmp))
        return tmp

    def scatter3(self):
        for names, rows, tmp in self._pending:
            for j, n in enumerate(names):
                part = tmp[j * rows:(j + 1) * rows]
                sink = self.sinks.get(n)
                if sink is not None:
                    sink.add_(part)
                    self[n] = sink
                else:
                    self[n] = part
        self._pending = []


class TowerTrainer(object):
    """Runs a TowerEngine's weights in training mode.  `engine.w` holds the 16-bit copies of the parameters."""

    def __init__(self, engine, p_hidden=0.0, p_attn=0.0, seed=None):
        self.e = engine
        self.p_hidden, self.p_attn = float(p_hidden), float(p_attn)
        if (self.p_hidden > 0 or self.p_attn > 0) and seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # torch's CPU generator: torch.manual_seed reproduces it
        self.seed = int(seed or 0)

    def _dropout(self, x, out, site, rows, cols, res=None):
        """out = dropout(x) (+ res) over the first `rows` rows (row pitch = x.stride(0) for all three)."""
        _lib.check(_lib.load().ldot_dropout(_lib.ptr(x), _lib.ptr(res), _lib.ptr(out), rows, cols, x.stride(0), self.p_hidden,
                                            self.seed, site, self.e.fmt, _lib.stream_ptr()))

    # ------------------------------------------------------------------------------------------------ primitives
    def _gemm(self, a, lda, a_mn, b, ldb, b_mn, out, ldo, M, N, K, bias=None, aux=None, ld_aux=0, epi=0, accumulate=False):
        lib = _lib.load()
        _lib.check(lib.ldot_gemm(_lib.ptr(a), lda, a_mn, _lib.ptr(b), ldb, b_mn, _lib.ptr(bias), _lib.ptr(aux), ld_aux,
                                 _lib.ptr(out), ldo, M, N, K, self.e.fmt, epi, int(out.dtype == torch.float32),
                                 int(accumulate), _lib.stream_ptr()))

    def _dgrad(self, dy, w, out, M, aux=None, epi=0, ld_dy=None, ldo=None):
        """out[M, in] = dy[M, out] . w[out, in] (+ aux | * GELU'(aux))"""
        O, I = w.shape
        self._gemm(dy, O if ld_dy is None else ld_dy, 0, w, I, 1, out, I if ldo is None else ldo, M, I, O, aux=aux,
                   ld_aux=0 if aux is None else aux.stride(0), epi=epi)

    def _wgrad(self, dy, x, dw, rows, ld_x=None):
        """dw[out, in] += dy[rows, out]^T . x[rows, in]"""
        O, I = dw.shape
        self._gemm(dy, O, 1, x, I if ld_x is None else ld_x, 1, dw, I, O, I, rows, accumulate=True)

    def _colsum(self, x, out, rows):
        lib = _lib.load()
        _lib.check(lib.ldot_colsum16(_lib.ptr(x), x.stride(0), rows, x.shape[1], _lib.ptr(out), self.e.fmt, _lib.stream_ptr()))

    def _ln_bwd(self, dy, x, gamma, dx, dgamma, dbeta, dxsum, rows, H, ld_dy=None, ld_x=None, ld_dx=None):
        lib = _lib.load()
        f32 = torch.float32
        _lib.check(lib.ldot_layernorm_bwd(
            _lib.ptr(dy), dy.stride(0) if ld_dy is None else ld_dy, int(dy.dtype == f32),
            _lib.ptr(x), x.stride(0) if ld_x is None else ld_x, int(x.dtype == f32), _lib.ptr(gamma),
            _lib.ptr(dx), dx.stride(0) if ld_dx is None else ld_dx, int(dx.dtype == f32),
            _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dxsum), rows, H, self.e.fmt, _lib.stream_ptr()))

    def _gelu(self, x, out):
        _lib.check(_lib.load().ldot_gelu(_lib.ptr(x), _lib.ptr(out), x.numel(), self.e.fmt, _lib.stream_ptr()))

    def _linear_dropout(self, a, wt, bias, residual, out, M, site):
        """out = dropout(a . wt^T + bias) + residual in ONE kernel (the mask of _dropout(site) in the GEMM epilogue)."""
        N, K = wt.shape
        _lib.check(_lib.load().ldot_linear_dropout(
            _lib.ptr(a), a.stride(0), _lib.ptr(wt), K, _lib.ptr(bias), _lib.ptr(residual), residual.stride(0),
            _lib.ptr(out), out.stride(0), M, N, K, self.e.fmt, self.p_hidden, self.seed, site, _lib.stream_ptr()))

    def _linear_gelu_grad(self, a, wt, bias, pre, out, M):
        """With z = a . wt^T + bias (rounded to 16 bit): out = GELU(z) and pre = GELU'(z) from one kernel."""
        N, K = wt.shape
        _lib.check(_lib.load().ldot_linear_gelu_grad(
            _lib.ptr(a), a.stride(0), _lib.ptr(wt), K, _lib.ptr(bias), _lib.ptr(pre), pre.stride(0), _lib.ptr(out),
            out.stride(0), M, N, K, self.e.fmt, _lib.stream_ptr()))

    def _ln_bwd_dropout(self, dy, x, gamma, dx, dx_masked, dgamma, dbeta, dxsum, rows, H, site):
        """LayerNorm backward whose input was dropout(dense) + residual: dx (residual branch), dx_masked (dense branch)
        and the dense bias gradient (column sums of dx_masked) in one pass."""
        _lib.check(_lib.load().ldot_layernorm_bwd_dropout(
            _lib.ptr(dy), dy.stride(0), _lib.ptr(x), x.stride(0), _lib.ptr(gamma), _lib.ptr(dx), _lib.ptr(dx_masked),
            dx.stride(0), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dxsum), rows, H, self.p_hidden, self.seed, site,
            self.e.fmt, _lib.stream_ptr()))

    # ------------------------------------------------------------------------------------------------ forward
    def _forward_layers(self, h, mask, B, S, tape):
        e, lib = self.e, _lib.load()
        T, H, F, dt, dev, w = B * S, e.H, e.ffn, e.dtype, h.device, e.w
        stream = _lib.stream_ptr()
        tape.layers = []
        tape.p_hidden, tape.p_attn, tape.seed = self.p_hidden, self.p_attn, self.seed
        ph = self.p_hidden > 0
        if ph:
            self._dropout(h, h, SITE_EMB, T, H)
        for i in range(e.layers):
            qkv = torch.empty((T, 3 * H), dtype=dt, device=dev)
            ctx = torch.empty((T, H), dtype=dt, device=dev)
            pre1 = torch.empty((T, H), dtype=dt, device=dev)
            a = torch.empty((T, H), dtype=dt, device=dev)
            fpre = torch.empty((T, F), dtype=dt, device=dev)
            f = torch.empty((T, F), dtype=dt, device=dev)
            pre2 = torch.empty((T, H), dtype=dt, device=dev)
            out = torch.empty((T, H), dtype=dt, device=dev)
            e._linear(h, H, w[f"qkv_w{i}"], w[f"qkv_b{i}"], qkv, T)
            _lib.check(lib.ldot_attention_train(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx), B, S, H, e.heads, self.p_attn,
                                                self.seed, 4 * i, e.fmt, stream))
            if ph:   # dense -> dropout -> + residual (layer.py:111-115), one kernel
                self._linear_dropout(ctx, w[f"o_w{i}"], w[f"o_b{i}"], h, pre1, T, 4 * i + 1)
            else:
                e._linear(ctx, H, w[f"o_w{i}"], w[f"o_b{i}"], pre1, T, residual=h)
            e._layernorm(pre1, w[f"ln1_g{i}"], w[f"ln1_b{i}"], a, T, H)
            # FFN-up: z = a W1^T + b1 from the plain GEMM, then one bandwidth-bound pass f = GELU(z), fpre <- GELU'(z)
            # (the epilogue that evaluates both - _linear_gelu_grad - is issue-bound at K = 768 and slower than the pair)
            e._linear(a, H, w[f"f1_w{i}"], w[f"f1_b{i}"], fpre, T)
            _lib.check(lib.ldot_gelu_grad(_lib.ptr(fpre), _lib.ptr(f), fpre.numel(), e.fmt, stream))
            if ph:
                self._linear_dropout(f, w[f"f2_w{i}"], w[f"f2_b{i}"], a, pre2, T, 4 * i + 2)
            else:
                e._linear(f, F, w[f"f2_w{i}"], w[f"f2_b{i}"], pre2, T, residual=a)
            e._layernorm(pre2, w[f"ln2_g{i}"], w[f"ln2_b{i}"], out, T, H)
            tape.layers.append((h, qkv, ctx, pre1, a, fpre, f, pre2))
            h = out
        tape.h_last = h
        return h

    def _forward_head(self, h, B, S, tape):
        e, w, H, dev, dt = self.e, self.e.w, self.e.H, h.device, self.e.dtype
        if not e.project:
            return h.view(B, S, H)[:, 0, :].float()
        x0 = torch.empty((B, 2 * H), dtype=dt, device=dev)
        x1 = torch.empty((B, 2 * H), dtype=dt, device=dev)
        x2 = torch.empty((B, 2 * H), dtype=dt, device=dev)
        out = torch.empty((B, e.out_dim), dtype=torch.float32, device=dev)
        e._linear(h, S * H, w["p0_w"], w["p0_b"], x0, B)
        self._gelu(x0, x1)
        e._layernorm(x1, w["p_ln_g"], w["p_ln_b"], x2, B, 2 * H)
        e._linear(x2, 2 * H, w["p3_w"], w["p3_b"], out, B)
        tape.x0, tape.x1, tape.x2 = x0, x1, x2
        return out

    def forward_text(self, input_ids, attention_mask, position_ids):
        e, dev = self.e, self.e.device
        ids, mask, pos = e._i64(input_ids, dev), e._i64(attention_mask, dev), e._i64(position_ids, dev)
        B, L = ids.shape
        if L > 128:
            raise ValueError(f"sequence length {L} > 128 is not supported by the attention kernel")
        if pos.dim() == 1:
            pos = pos[None, :]
        tape = _Tape()
        tape.kind, tape.B, tape.S, tape.Lt, tape.R = "txt", B, L, L, 0
        tape.ids, tape.pos, tape.mask = ids, pos, mask
        h = torch.empty((B * L, e.H), dtype=e.dtype, device=dev)
        e._embed_text(ids, pos, h, B, L, L)
        h = self._forward_layers(h, mask, B, L, tape)
        return self._forward_head(h, B, L, tape), tape

    def forward_image(self, input_ids, attention_mask, position_ids, img_feat, img_pos_feat, gather_index=None):
        e, dev, w, H, lib = self.e, self.e.device, self.e.w, self.e.H, _lib.load()
        ids, mask, pos = e._i64(input_ids, dev), e._i64(attention_mask, dev), e._i64(position_ids, dev)
        feat = img_feat.to(device=dev, dtype=torch.float32).contiguous()
        box = img_pos_feat.to(device=dev, dtype=torch.float32).contiguous()
        B, Lt = ids.shape
        R = feat.shape[1]
        S = Lt + R
        if S > 128:
            raise ValueError(f"sequence length {S} > 128 is not supported by the attention kernel")
        if mask.shape[1] != S:
            raise ValueError(f"attention_mask has {mask.shape[1]} positions, expected {S}")
        if gather_index is not None and not torch.cuda.is_current_stream_capturing():
            # (a captured step was run eagerly on the same static batch first: checked there)
            gi = gather_index.to(dev)
            if not torch.equal(gi, torch.arange(S, device=dev)[None, :].expand(B, S)):
                raise NotImplementedError("only the identity gather_index of dvl/data/itm.py is supported")
        if pos.dim() == 1:
            pos = pos[None, :]
        stream = _lib.stream_ptr()
        tape = _Tape()
        tape.kind, tape.B, tape.S, tape.Lt, tape.R = "img", B, S, Lt, R
        tape.ids, tape.pos, tape.mask, tape.box = ids, pos, mask, box
        h = torch.empty((B * S, H), dtype=e.dtype, device=dev)
        e._embed_text(ids, pos, h, B, Lt, S)
        f16 = torch.empty((B * R, e.img_dim), dtype=e.dtype, device=dev)
        _lib.check(lib.ldot_cast_f32(_lib.ptr(feat), _lib.ptr(f16), B * R * e.img_dim, e.fmt, stream))
        lin = torch.empty((B * R, H), dtype=torch.float32, device=dev)
        e._linear(f16, e.img_dim, w["img_w"], w["img_bias"], lin, B * R)
        _lib.check(lib.ldot_embed_image(
            _lib.ptr(lin), _lib.ptr(box), _lib.ptr(w["img_ln_g"]), _lib.ptr(w["img_ln_b"]), _lib.ptr(w["pos_w"]),
            _lib.ptr(w["pos_bias"]), _lib.ptr(w["pos_ln_g"]), _lib.ptr(w["pos_ln_b"]), _lib.ptr(w["type1_f32"]),
            _lib.ptr(w["iemb_ln_g"]), _lib.ptr(w["iemb_ln_b"]), _lib.ptr(h), B, R, S, Lt, H, e.fmt, stream))
        tape.feat16, tape.lin = f16, lin
        h = self._forward_layers(h, mask, B, S, tape)
        return self._forward_head(h, B, S, tape), tape

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, tape, d_pooled, sinks=None, done=None):
        """d_pooled fp32 [B, D] -> {reference parameter name: fp32 gradient}.  `sinks`: {name: existing fp32 gradient
        tensor}: those gradients are ACCUMULATED into the given tensors (and still listed in the result).
        `done(names)`: called as soon as the gradients of `names` are final (head, then each layer from the last to the
        first, then the embeddings) - what lets the gradient all-reduce start under the rest of the backward."""
        e, w, lib = self.e, self.e.w, _lib.load()
        H, F, dt, dev = e.H, e.ffn, e.dtype, d_pooled.device
        B, S = tape.B, tape.S
        T = B * S
        stream = _lib.stream_ptr()
        f32 = torch.float32
        sinks = sinks or {}
        g = _GradSinks(sinks, dev)

        def buf(rows, cols):
            return torch.empty((rows, cols), dtype=dt, device=dev)

        d_h = torch.zeros((T, H), dtype=dt, device=dev)
        d_pooled = d_pooled.to(f32).contiguous()
        if e.project:
            D = e.out_dim
            dO = buf(B, D)
            _lib.check(lib.ldot_cast_f32(_lib.ptr(d_pooled), _lib.ptr(dO), B * D, e.fmt, stream))
            g.z("encode_proj.3.weight", D, 2 * H)
            g.z("encode_proj.3.bias", D)
            self._wgrad(dO, tape.x2, g["encode_proj.3.weight"], B)
            self._colsum(dO, g["encode_proj.3.bias"], B)
            d_x2 = buf(B, 2 * H)
            self._dgrad(dO, w["p3_w"], d_x2, B)
            d_x1 = buf(B, 2 * H)
            g.z("encode_proj.2.weight", 2 * H)
            g.z("encode_proj.2.bias", 2 * H)
            self._ln_bwd(d_x2, tape.x1, w["p_ln_g"], d_x1, g["encode_proj.2.weight"], g["encode_proj.2.bias"], None, B, 2 * H)
            d_x0 = buf(B, 2 * H)
            _lib.check(lib.ldot_gelu_bwd(_lib.ptr(tape.x0), _lib.ptr(d_x1), _lib.ptr(d_x0), B * 2 * H, e.fmt, stream))
            g.z("encode_proj.0.weight", 2 * H, H)
            g.z("encode_proj.0.bias", 2 * H)
            self._wgrad(d_x0, tape.h_last, g["encode_proj.0.weight"], B, ld_x=S * H)
            self._colsum(d_x0, g["encode_proj.0.bias"], B)
            self._dgrad(d_x0, w["p0_w"], d_h, B, ldo=S * H)
        else:
            d_h.view(B, S, H)[:, 0, :] = d_pooled.to(dt)

        ph = tape.p_hidden > 0
        self.p_hidden, self.p_attn, self.seed = tape.p_hidden, tape.p_attn, tape.seed   # (the forward's masks)
        reported = set()
        for i in reversed(range(e.layers)):
            x, qkv, ctx, pre1, a, fpre, f, pre2 = tape.layers[i]
            p = f"bert.encoder.layer.{i}."
            # BertOutput: h = LN(f W2^T + b2 + a)
            d_pre2 = buf(T, H)
            g.z(p + "output.LayerNorm.weight", H)
            g.z(p + "output.LayerNorm.bias", H)
            g.z(p + "output.dense.bias", H)
            # (with hidden dropout the dense output's gradient is the MASKED d_pre2 and its bias gradient the column sums
            # of that; the residual branch keeps the unmasked one - the fused kernel writes both)
            if ph:
                d_o2 = buf(T, H)
                self._ln_bwd_dropout(d_h, pre2, w[f"ln2_g{i}"], d_pre2, d_o2, g[p + "output.LayerNorm.weight"],
                                     g[p + "output.LayerNorm.bias"], g[p + "output.dense.bias"], T, H, 4 * i + 2)
            else:
                self._ln_bwd(d_h, pre2, w[f"ln2_g{i}"], d_pre2, g[p + "output.LayerNorm.weight"],
                             g[p + "output.LayerNorm.bias"], g[p + "output.dense.bias"], T, H)
                d_o2 = d_pre2
            g.z(p + "output.dense.weight", H, F)
            self._wgrad(d_o2, f, g[p + "output.dense.weight"], T)
            d_fpre = buf(T, F)
            self._dgrad(d_o2, w[f"f2_w{i}"], d_fpre, T, aux=fpre, epi=4)   # d z = (d_o2 W2) * GELU'(z)
            del d_o2
            # BertIntermediate
            g.z(p + "intermediate.dense.weight", F, H)
            g.z(p + "intermediate.dense.bias", F)
            self._wgrad(d_fpre, a, g[p + "intermediate.dense.weight"], T)
            self._colsum(d_fpre, g[p + "intermediate.dense.bias"], T)
            d_a = buf(T, H)
            self._dgrad(d_fpre, w[f"f1_w{i}"], d_a, T, aux=d_pre2, epi=3)
            del d_fpre
            # BertSelfOutput: a = LN(ctx Wo^T + bo + x)
            d_pre1 = buf(T, H)
            g.z(p + "attention.output.LayerNorm.weight", H)
            g.z(p + "attention.output.LayerNorm.bias", H)
            g.z(p + "attention.output.dense.bias", H)
            if ph:
                d_o1 = buf(T, H)
                self._ln_bwd_dropout(d_a, pre1, w[f"ln1_g{i}"], d_pre1, d_o1, g[p + "attention.output.LayerNorm.weight"],
                                     g[p + "attention.output.LayerNorm.bias"], g[p + "attention.output.dense.bias"], T, H,
                                     4 * i + 1)
            else:
                self._ln_bwd(d_a, pre1, w[f"ln1_g{i}"], d_pre1, g[p + "attention.output.LayerNorm.weight"],
                             g[p + "attention.output.LayerNorm.bias"], g[p + "attention.output.dense.bias"], T, H)
                d_o1 = d_pre1
            g.z(p + "attention.output.dense.weight", H, H)
            self._wgrad(d_o1, ctx, g[p + "attention.output.dense.weight"], T)
            d_ctx = buf(T, H)
            self._dgrad(d_o1, w[f"o_w{i}"], d_ctx, T)
            del d_o1
            # self-attention
            d_qkv = buf(T, 3 * H)
            _lib.check(lib.ldot_attention_bwd(_lib.ptr(qkv), _lib.ptr(tape.mask), _lib.ptr(ctx), _lib.ptr(d_ctx),
                                              _lib.ptr(d_qkv), B, S, H, e.heads, tape.p_attn, tape.seed, 4 * i, e.fmt,
                                              stream))
            dwqkv = g.z3([p + f"attention.self.{nm}.weight" for nm in ("query", "key", "value")], H, H)
            dbqkv = g.z3([p + f"attention.self.{nm}.bias" for nm in ("query", "key", "value")], H)
            self._wgrad(d_qkv, x, dwqkv, T)
            self._colsum(d_qkv, dbqkv, T)
            g.scatter3()
            d_x = buf(T, H)
            self._dgrad(d_qkv, w[f"qkv_w{i}"], d_x, T, aux=d_pre1, epi=3)
            d_h = d_x
            tape.layers[i] = None   # release this layer's activations
            if done is not None:
                done([n for n in g if n not in reported])
                reported.update(g)

        if ph:
            self._dropout(d_h, d_h, SITE_EMB, T, H)
        self._backward_embeddings(tape, d_h, g)
        if done is not None:
            done([n for n in g if n not in reported])
        return g

    def _backward_embeddings(self, tape, d_h, g):
        e, w, lib = self.e, self.e.w, _lib.load()
        H, dev, f32 = e.H, d_h.device, torch.float32
        B, S, Lt, R = tape.B, tape.S, tape.Lt, tape.R
        stream = _lib.stream_ptr()

        pe = "bert.embeddings."
        g.z(pe + "word_embeddings.weight", e.vocab, H)
        g.z(pe + "position_embeddings.weight", e.max_pos, H)
        g.z(pe + "token_type_embeddings.weight", 2, H)
        g.z(pe + "LayerNorm.weight", H)
        g.z(pe + "LayerNorm.bias", H)
        # text positions: rows b * S + l, l < Lt
        d_txt = d_h if Lt == S else d_h.view(B, S, H)[:, :Lt, :].contiguous().view(B * Lt, H)
        pos_stride = 0 if tape.pos.shape[0] == 1 else tape.pos.stride(0)
        s = torch.empty((B * Lt, H), dtype=f32, device=dev)
        _lib.check(lib.ldot_embed_text_sum(_lib.ptr(tape.ids), _lib.ptr(tape.pos), pos_stride, _lib.ptr(w["word"]),
                                           _lib.ptr(w["pos"]), _lib.ptr(w["type0"]), _lib.ptr(s), B, Lt, H, e.vocab,
                                           e.max_pos, e.fmt, stream))
        dxe = torch.empty((B * Lt, H), dtype=f32, device=dev)
        self._ln_bwd(d_txt, s, w["emb_ln_g"], dxe, g[pe + "LayerNorm.weight"], g[pe + "LayerNorm.bias"],
                     g[pe + "token_type_embeddings.weight"][0], B * Lt, H)
        _lib.check(lib.ldot_embed_scatter(_lib.ptr(dxe), _lib.ptr(tape.ids), _lib.ptr(tape.pos), pos_stride,
                                          _lib.ptr(g[pe + "word_embeddings.weight"]),
                                          _lib.ptr(g[pe + "position_embeddings.weight"]), B, Lt, H, e.vocab, e.max_pos,
                                          stream))
        if tape.kind != "img":
            return
        pi = "bert.img_embeddings."
        rows = B * R
        d_img = d_h.view(B, S, H)[:, Lt:, :].contiguous().view(rows, H)
        q = torch.empty((rows, H), dtype=f32, device=dev)
        spre = torch.empty((rows, H), dtype=f32, device=dev)
        _lib.check(lib.ldot_embed_image_pre(_lib.ptr(tape.lin), _lib.ptr(tape.box), _lib.ptr(w["img_ln_g"]),
                                            _lib.ptr(w["img_ln_b"]), _lib.ptr(w["pos_w"]), _lib.ptr(w["pos_bias"]),
                                            _lib.ptr(w["pos_ln_g"]), _lib.ptr(w["pos_ln_b"]), _lib.ptr(w["type1_f32"]),
                                            _lib.ptr(q), _lib.ptr(spre), rows, H, stream))
        for nm in ("LayerNorm", "img_layer_norm", "pos_layer_norm"):
            g.z(pi + nm + ".weight", H)
            g.z(pi + nm + ".bias", H)
        g.z(pi + "img_linear.bias", H)
        g.z(pi + "pos_linear.bias", H)
        ds = torch.empty((rows, H), dtype=f32, device=dev)
        self._ln_bwd(d_img, spre, w["iemb_ln_g"], ds, g[pi + "LayerNorm.weight"], g[pi + "LayerNorm.bias"],
                     g[pe + "token_type_embeddings.weight"][1], rows, H)
        d_lin = torch.empty((rows, H), dtype=e.dtype, device=dev)
        self._ln_bwd(ds, tape.lin, w["img_ln_g"], d_lin, g[pi + "img_layer_norm.weight"], g[pi + "img_layer_norm.bias"],
                     g[pi + "img_linear.bias"], rows, H)
        dq = torch.empty((rows, H), dtype=f32, device=dev)
        self._ln_bwd(ds, q, w["pos_ln_g"], dq, g[pi + "pos_layer_norm.weight"], g[pi + "pos_layer_norm.bias"],
                     g[pi + "pos_linear.bias"], rows, H)
        g.z(pi + "pos_linear.weight", H, 7)
        _lib.check(lib.ldot_pos_wgrad(_lib.ptr(dq), _lib.ptr(tape.box), rows, H, _lib.ptr(g[pi + "pos_linear.weight"]), stream))
        g.z(pi + "img_linear.weight", H, e.img_dim)
        self._wgrad(d_lin, tape.feat16, g[pi + "img_linear.weight"], rows)


# ------------------------------------------------------------------------------------------------------- autograd
class TowerFunction(torch.autograd.Function):
    """pooled = tower(inputs; parameters): one autograd node per tower call.  `names` are the reference parameter
    names of `params` (same order); parameters the forward never reads (pooler, mask_embedding) are not passed."""

    @staticmethod
    def forward(ctx, runner, names, *params):
        pooled, tape = runner()
        ctx.tape, ctx.names = tape, names
        ctx.trainer = runner.trainer
        ctx.params = params
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        if ctx.tape is None:
            raise RuntimeError("the tower's activations were already released: backward through a tower call runs once")
        # Parameters that already own an fp32 .grad (FusedAdamW's flat views, or a previous backward) are accumulated
        # INTO by the kernels (every parameter-gradient kernel adds), and autograd is told there is nothing to add:
        # no zero-filled temporaries and no per-parameter add launches.
        sinks = {}
        for i, (nm, p) in enumerate(zip(ctx.names, ctx.params)):
            g = p.grad
            if ctx.needs_input_grad[2 + i] and g is not None and g.dtype == torch.float32 and g.is_contiguous() \
                    and g.device == d_pooled.device and g.shape == p.shape:
                sinks[nm] = g
        grads = ctx.trainer.backward(ctx.tape, d_pooled, sinks, done=_early_sync_hook(ctx.names, ctx.params, sinks))
        ctx.tape = None
        out = []
        for i, nm in enumerate(ctx.names):
            out.append(grads.get(nm) if (ctx.needs_input_grad[2 + i] and nm not in sinks) else None)
        return (None, None) + tuple(out)


def _early_sync_hook(names, params, sinks):
    """-> done(names) for TowerTrainer.backward, or None: hands finished gradients to the FusedAdamW optimisers that asked
    for an overlapped gradient all-reduce (`overlap_grad_sync`) and own the flat buffers those gradients live in."""
    owners = []
    for f in _lib.flat_buffers:
        opt = f.get("owner") and f["owner"]()
        if opt is not None and opt.overlap_grad_sync and opt.distributed and opt not in owners:
            owners.append(opt)
    if not owners:
        return None
    by_name = dict(zip(names, params))

    def done(finished):
        ps = [by_name[n] for n in finished if n in sinks and n in by_name]
        if ps:
            for opt in owners:
                opt.reduce_early(ps)
    return done


class _Runner(object):
    def __init__(self, trainer, fn):
        self.trainer, self.fn = trainer, fn

    def __call__(self):
        return self.fn()


UNUSED_PARAMETERS = ("bert.pooler.", "bert.img_embeddings.mask_embedding.")


def run_tower_training(module, engine, kind, inputs):
    """module: BertEncoder / UniterEncoder (parameters under the reference names); returns pooled fp32 [B, D] attached
    to the autograd graph.  Dropout as the module's mode says: train() -> config probabilities, eval() -> none;
    `module.dropout_seed` (tests) pins the mask seed of the next calls."""
    c = module.config
    on = module.training
    trainer = TowerTrainer(engine, c.hidden_dropout_prob if on else 0.0, c.attention_probs_dropout_prob if on else 0.0,
                           getattr(module, "dropout_seed", None))
    named = [(n, p) for n, p in module.named_parameters() if not n.startswith(UNUSED_PARAMETERS)]
    names = tuple(n for n, _ in named)
    if kind == "txt":
        runner = _Runner(trainer, lambda: trainer.forward_text(*inputs))
    else:
        runner = _Runner(trainer, lambda: trainer.forward_image(*inputs))
    return TowerFunction.apply(runner, names, *[p for _, p in named])


class NllFunction(torch.autograd.Function):
    """loss = in-batch NLL(q, ctx[, captions]) with gradients to the embeddings (bi_encoder.py:615-656 under autograd).
    Forward: hi/lo-split tcgen05 scores + fused row log-sum-exp.  Backward: d scores (bf16) from one kernel, then
    dQ = dS . C (dgrad form) and dC = dS^T . Q (wgrad form) on the tensor cores."""

    @staticmethod
    def forward(ctx, q, c, cap, pos, cap_weight, reduction, scores_fn):
        lib = _lib.load()
        use_cap = cap is not None and cap_weight != 0
        s1 = scores_fn(q, c).contiguous()
        s2 = scores_fn(q, cap).contiguous() if use_cap else None
        bq, bc = s1.shape
        dev = s1.device
        scores = torch.empty((bq, bc), dtype=torch.float32, device=dev)
        row_loss = torch.empty((bq,), dtype=torch.float32, device=dev)
        row_correct = torch.empty((bq,), dtype=torch.int32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        correct = torch.empty((), dtype=torch.int64, device=dev)
        _lib.check(lib.ldot_inbatch_nll(_lib.ptr(s1), _lib.ptr(s2), float(cap_weight), _lib.ptr(pos), bq, bc, reduction,
                                        _lib.ptr(scores), _lib.ptr(row_loss), _lib.ptr(row_correct), _lib.ptr(loss),
                                        _lib.ptr(correct), _lib.stream_ptr()))
        ctx.save_for_backward(q, c, cap if use_cap else None, pos, scores)
        ctx.use_cap, ctx.cap_weight, ctx.reduction = use_cap, float(cap_weight), reduction
        ctx.mark_non_differentiable(correct, scores, s1)
        return loss, correct, scores, s1

    @staticmethod
    def backward(ctx, d_loss, _dc, _ds, _ds1):
        lib = _lib.load()
        q, c, cap, pos, scores = ctx.saved_tensors
        bq, bc = scores.shape
        D = q.shape[1]
        dev = q.device
        stream = _lib.stream_ptr()
        fmt, dt = _lib.COARSE_BF16, torch.bfloat16
        ld = (bc + 7) // 8 * 8
        bqp = (bq + 7) // 8 * 8
        ds = torch.empty((bq, ld), dtype=dt, device=dev)
        up = d_loss.to(torch.float32).reshape(1).contiguous()
        _lib.check(lib.ldot_inbatch_nll_bwd(_lib.ptr(scores), _lib.ptr(pos), bq, bc, _lib.ptr(up), ctx.reduction,
                                            _lib.ptr(ds), ld, fmt, stream))

        def cast16(x, rows_padded):
            out = torch.zeros((rows_padded, D), dtype=dt, device=dev)
            xf = x.detach().to(torch.float32).contiguous()
            _lib.check(lib.ldot_cast_f32(_lib.ptr(xf), _lib.ptr(out), x.shape[0] * D, fmt, stream))
            return out

        def gemm(a, lda, a_mn, b, ldb, out, ldo, M, N, K, acc):
            _lib.check(lib.ldot_gemm(_lib.ptr(a), lda, a_mn, _lib.ptr(b), ldb, 1, None, None, 0, _lib.ptr(out), ldo, M, N, K,
                                     fmt, 0, 1, int(acc), stream))

        need_q, need_c, need_cap = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and ctx.use_cap
        w = ctx.cap_weight if ctx.use_cap else 0.0
        dq = dc = dcap = None
        if need_q:
            dq = torch.empty((bq, D), dtype=torch.float32, device=dev)
            mix = c if not ctx.use_cap else (1.0 - w) * c.detach().float() + w * cap.detach().float()
            gemm(ds, ld, 0, cast16(mix, ld), D, dq, D, bq, D, ld, False)       # dQ = dS . ((1 - w) C + w Cap)
        if need_c or need_cap:
            dct = torch.zeros((ld, D), dtype=torch.float32, device=dev)
            gemm(ds, ld, 1, cast16(q, bqp)[:bq], D, dct, D, ld, D, bq, True)   # dS^T . Q
            if need_c:
                dc = dct[:bc] * (1.0 - w) if ctx.use_cap else dct[:bc]
            if need_cap:
                dcap = dct[:bc] * w
        return dq, dc, dcap, None, None, None, None


# ------------------------------------------------------------------------------------------------------- optimiser
def zero_grads(module, set_to_none=True):
    """module.zero_grad() that keeps flat-buffer gradients attached: a flat gradient buffer all of whose parameters
    belong to `module` is cleared with one memset, other flat views are zeroed in place, everything else follows
    nn.Module.zero_grad."""
    params = list(module.parameters())
    mine = set(id(p) for p in params)
    handled = set()
    for f in _lib.flat_buffers:
        if all(id(p) in mine for p in f["params"]):
            f["g"].zero_()
            base, n = f["g"].data_ptr(), f["g"].numel()
            for p in f["params"]:
                if p.grad is not None and base <= p.grad.data_ptr() < base + 4 * n:
                    handled.add(id(p))
    spans = [(f["g"].data_ptr(), f["g"].data_ptr() + 4 * f["g"].numel()) for f in _lib.flat_buffers]
    for p in params:
        if p.grad is None or id(p) in handled:
            continue
        ptr = p.grad.data_ptr()
        if any(lo <= ptr < hi for lo, hi in spans):
            p.grad.zero_()
        elif set_to_none:
            p.grad = None
        else:
            p.grad.detach_()
            p.grad.zero_()


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics - decoupled weight decay applied to the parameter BEFORE the Adam update, eps added
    after the bias-corrected sqrt(v), one step counter for all parameters - which is what bi_encoder.py:566-576's
    transformers.AdamW(correct_bias=True) computes up to the placement of eps (HF adds it before the bias-2 correction) and
    the order of the decay term (HF decays after the update); with the shipped eps = 1e-8 / weight_decay = 0 the two rules
    differ by O(eps) per step.  ONE kernel launch per parameter group: parameters, gradients and both moments of a group
    live in flat fp32 buffers (the nn.Parameters become views of the flat master copy), `step()` is ldot_adamw over the
    flat buffers, optionally preceded by the global gradient-norm clip of train_itm.py:262-267 (`max_grad_norm` > 0:
    one extra reduction launch per group instead of a pass per tensor).

    Like torch, parameters that have never received a gradient (the unused pooler, mask_embedding) are left alone:
    the flat buffers are laid out at the first step() from the parameters that carry a gradient, and re-laid out if
    another parameter starts receiving one.  `state_dict()` has torch.optim.AdamW's layout (step / exp_avg /
    exp_avg_sq per parameter)."""

    def __init__(self, params, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0,
                 shadow_dtype=torch.bfloat16):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        # 16-bit mirror of every group that holds matrices, written by the AdamW kernel: the towers alias their GEMM
        # operands to it (towers.py: load), so a step needs no weight re-conversion.  None disables the mirror.
        self.shadow_dtype = shadow_dtype
        self.distributed = False      # set by setup_for_distributed_mode: average gradients over the ranks in step()
        self._synced = False
        self._flat = None
        self._steps = 0
        # {group index: fp32 [3] device tensor} once device_hyper(True) was called: lr and the bias corrections are then
        # read by the kernel from device memory (ldot_adamw_dev), which is what lets step() be captured in a CUDA graph
        self._hyper = None
        # Overlapped gradient averaging (opt-in: valid when every step() follows ONE backward - no gradient accumulation):
        # the towers report finished gradients during backward (training._early_sync_hook) and their spans of the flat
        # buffers are all-reduced asynchronously, under the rest of the backward; sync_gradients() reduces what is left.
        self.overlap_grad_sync = False
        self.early_sync_bytes = 32 << 20      # spans are batched up to this size per collective
        self._early_pending = []              # parameters whose gradients are final, not yet handed to NCCL
        self._early_spans = []                # (flat buffer index, lo, hi) already being reduced
        self._early_works = []

    PAD = 64   # parameters start on 64-element boundaries (TMA operands need 16-byte aligned bases)

    def _flatten(self):
        flat = []
        self._unregister()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad and (p.grad is not None or "exp_avg" in self.state.get(p, {}))]
            if not ps:
                flat.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise _lib.LdotError("FusedAdamW runs only on CUDA parameters (move the model to the GPU first)")
            pad = self.PAD
            n = sum((p.numel() + pad - 1) // pad * pad for p in ps)
            f = dict(p=torch.zeros(n, dtype=torch.float32, device=dev), g=torch.zeros(n, dtype=torch.float32, device=dev),
                     m=torch.zeros(n, dtype=torch.float32, device=dev), v=torch.zeros(n, dtype=torch.float32, device=dev),
                     params=ps, ids=set(id(p) for p in ps), offsets=[], p16=None)
            off = 0
            for p in ps:
                k = p.numel()
                f["offsets"].append(off)
                sl = slice(off, off + k)
                f["p"][sl].copy_(p.data.reshape(-1))
                p.data = f["p"][sl].view(p.shape)
                if p.grad is not None:
                    f["g"][sl].copy_(p.grad.reshape(-1))
                p.grad = f["g"][sl].view(p.shape)
                old = self.state.get(p, {})
                if "exp_avg" in old:
                    f["m"][sl].copy_(old["exp_avg"].reshape(-1))
                    f["v"][sl].copy_(old["exp_avg_sq"].reshape(-1))
                    self._steps = max(self._steps, int(old.get("step", 0)))
                self.state[p] = {"step": torch.tensor(float(self._steps)), "exp_avg": f["m"][sl].view(p.shape),
                                 "exp_avg_sq": f["v"][sl].view(p.shape)}
                off += (k + pad - 1) // pad * pad
            if self.shadow_dtype is not None and any(p.dim() >= 2 for p in ps):
                f["p16"] = f["p"].to(self.shadow_dtype)
            f["owner"] = weakref.ref(self)
            flat.append(f)
        self._flat = flat
        _lib.flat_buffers.extend(f for f in flat if f is not None)
        _lib.param_generation[0] += 1

    def _relayout(self):
        """Drop the flat buffers (the mirror dtype changed): the next step() lays them out again, adopting the moments."""
        self._unregister()
        return None

    def _unregister(self):
        if self._flat:
            mine = set(id(f) for f in self._flat if f is not None)
            _lib.flat_buffers[:] = [f for f in _lib.flat_buffers if id(f) not in mine]

    def __del__(self):
        try:
            self._unregister()
        except Exception:
            pass

    def state_dict(self):
        for f in self._flat or []:
            if f is not None:
                for p in f["params"]:
                    self.state[p]["step"].fill_(float(self._steps))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._unregister()    # (the old flat buffers must not stay listed in _lib.flat_buffers)
        self._flat = None     # re-laid out at the next step(), adopting the loaded moments and step count
        self._steps = 0

    def zero_grad(self, set_to_none=False):
        """Clears the flat gradient buffers in place (the .grad views stay attached)."""
        self._synced = False
        self._early_pending = []
        if self._early_works or self._early_spans:   # (a backward whose step() was skipped: its early reductions are void)
            for w in self._early_works:
                if w is not None:
                    w.wait()
            self._early_spans, self._early_works = [], []
        if self._flat is None:
            return super().zero_grad(set_to_none=True)
        for f in self._flat:
            if f is None:
                continue
            f["g"].zero_()
            for p, off in zip(f["params"], f["offsets"]):   # (model.zero_grad() may have dropped the views)
                k = p.numel()
                if p.grad is None or p.grad.data_ptr() != f["g"].data_ptr() + off * 4:
                    p.grad = f["g"][off:off + k].view(p.shape)

    def _needs_layout(self):
        if self._flat is None:
            return True
        for group, f in zip(self.param_groups, self._flat):
            known = f["ids"] if f is not None else ()
            for p in group["params"]:
                if p.requires_grad and p.grad is not None and id(p) not in known:
                    return True
        return False

    def _collect(self):
        """Lay the flat buffers out if needed and pull in gradients that autograd left outside them (after
        model.zero_grad() dropped the views).  -> the live groups' buffers."""
        if self._needs_layout():
            self._flatten()
        live = [f for f in self._flat if f is not None]
        for f in live:
            for p, off in zip(f["params"], f["offsets"]):
                k = p.numel()
                if p.grad is None:
                    f["g"][off:off + k].zero_()
                    p.grad = f["g"][off:off + k].view(p.shape)
                elif p.grad.data_ptr() != f["g"].data_ptr() + off * 4:
                    f["g"][off:off + k].copy_(p.grad.reshape(-1))
                    p.grad = f["g"][off:off + k].view(p.shape)
        return live

    @torch.no_grad()
    def reduce_early(self, params, flush=False):
        """Gradients of `params` are final for this step: queue them, and once early_sync_bytes are pending (or `flush`)
        start the asynchronous average of their spans of the flat gradient buffers.  Every rank runs the same backward, so
        every rank issues the same collectives in the same order."""
        import torch.distributed as dist
        if self._flat is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if self._needs_layout():      # (a parameter is about to join the flat buffers: this step is reduced the plain way)
            return
        self._early_pending.extend(params)
        if not flush and sum(p.numel() for p in self._early_pending) * 4 < self.early_sync_bytes:
            return
        pending = set(id(p) for p in self._early_pending)
        self._early_pending = []
        pad = self.PAD
        for fi, f in enumerate(self._flat):
            if f is None:
                continue
            runs = []   # maximal runs of adjacent pending parameters: [lo, hi) in elements
            for p, off in zip(f["params"], f["offsets"]):
                if id(p) not in pending or p.grad is None or p.grad.data_ptr() != f["g"].data_ptr() + off * 4:
                    continue
                end = off + (p.numel() + pad - 1) // pad * pad
                if runs and runs[-1][1] == off:
                    runs[-1][1] = end
                else:
                    runs.append([off, end])
            for lo, hi in runs:
                hi = min(hi, f["g"].numel())
                self._early_works.append(dist.all_reduce(f["g"][lo:hi], op=dist.ReduceOp.AVG, async_op=True)
                                         if dist.get_backend() == "nccl" else
                                         self._gloo_avg(f["g"][lo:hi]))
                self._early_spans.append((fi, lo, hi))

    @staticmethod
    def _gloo_avg(t):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(dist.get_world_size())
        return None

    @torch.no_grad()
    def sync_gradients(self, group=None):
        """Average the gradients over the ranks: ONE all-reduce per parameter group over the flat buffer (minus the spans
        reduce_early already handled).  step() calls it when `distributed` is set; call it yourself after backward() when
        something (clip_grad_norm_, logging) must see the synchronised gradients before step()."""
        from .utils import sync_gradients
        if self._synced:
            return
        live = self._collect()
        if self._early_pending:
            self.reduce_early([], flush=True)
        if not self._early_spans:
            sync_gradients([f["g"] for f in live], group)
        else:
            rest = []
            for fi, f in enumerate(self._flat):
                if f is None:
                    continue
                pos, n = 0, f["g"].numel()
                for _, lo, hi in sorted(sp for sp in self._early_spans if sp[0] == fi):
                    if lo > pos:
                        rest.append(f["g"][pos:lo])
                    pos = max(pos, hi)
                if pos < n:
                    rest.append(f["g"][pos:n])
            sync_gradients(rest, group)
            for w in self._early_works:
                if w is not None:
                    w.wait()        # (stream-level: the compute stream waits for NCCL's, the host does not)
        self._early_spans, self._early_works = [], []
        self._synced = True

    def device_hyper(self, on=True):
        """Route lr / bias corrections through device memory (GraphedTrainStep); off restores by-value arguments."""
        self._hyper = {} if on else None

    def begin_step(self):
        """Host half of a step: count it and (device_hyper mode) upload this step's lr and bias corrections.  step() calls
        it; a replayed CUDA graph calls it itself before every replay, since step() only ran at capture time."""
        self._steps += 1
        if self._hyper is None:
            return
        for gi, (group, f) in enumerate(zip(self.param_groups, self._flat or [])):
            if f is None:
                continue
            # (the same arithmetic as ldot_adamw, which receives the betas as C floats: both entry points step alike)
            b1, b2 = (float(torch.tensor(b, dtype=torch.float32)) for b in group["betas"])
            host = torch.tensor([float(group["lr"]), 1.0 - b1 ** self._steps, (1.0 - b2 ** self._steps) ** 0.5],
                                dtype=torch.float32).pin_memory()
            if gi not in self._hyper:
                self._hyper[gi] = torch.empty(3, dtype=torch.float32, device=f["p"].device)
            self._hyper[gi].copy_(host, non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.load()
        live = self._collect()
        capturing = bool(live) and live[0]["p"].is_cuda and torch.cuda.is_current_stream_capturing()
        if not capturing:   # (a capture records launches without running them: nothing is stepped)
            self.begin_step()
        stream = _lib.stream_ptr()
        if not live:
            return None
        if self.distributed:
            self.sync_gradients()
        self._synced = False
        ss = None
        if self.max_grad_norm > 0:
            ss = torch.zeros(1, dtype=torch.float32, device=live[0]["p"].device)
            for f in live:
                _lib.check(lib.ldot_sumsq(_lib.ptr(f["g"]), f["g"].numel(), _lib.ptr(ss), stream))
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            b1, b2 = group["betas"]
            fmt = _lib.COARSE_FP16 if (f["p16"] is not None and f["p16"].dtype == torch.float16) else _lib.COARSE_BF16
            if self._hyper is not None:
                gi = next(j for j, gr in enumerate(self.param_groups) if gr is group)
                if gi not in self._hyper:
                    if capturing:
                        raise _lib.LdotError("FusedAdamW: run one eager step() in device_hyper mode before capturing")
                    self._steps -= 1
                    self.begin_step()
                _lib.check(lib.ldot_adamw_dev(_lib.ptr(f["p"]), _lib.ptr(f["g"]), _lib.ptr(f["m"]), _lib.ptr(f["v"]),
                                              _lib.ptr(f["p16"]), f["p"].numel(), _lib.ptr(self._hyper[gi]), float(b1),
                                              float(b2), float(group["eps"]), float(group["weight_decay"]), _lib.ptr(ss),
                                              self.max_grad_norm, fmt, stream))
                continue
            _lib.check(lib.ldot_adamw(_lib.ptr(f["p"]), _lib.ptr(f["g"]), _lib.ptr(f["m"]), _lib.ptr(f["v"]),
                                      _lib.ptr(f["p16"]), f["p"].numel(), float(group["lr"]), float(b1), float(b2),
                                      float(group["eps"]), float(group["weight_decay"]), self._steps, _lib.ptr(ss),
                                      self.max_grad_norm, fmt, stream))
        _lib.param_generation[0] += 1   # towers that hold converted COPIES of the weights (not aliases) are stale now
        return None


# ------------------------------------------------------------------------------------------------------- CUDA graph
def _map_leaves(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map_leaves(v, fn) for k, v in obj.items()}
    return obj


def _copy_leaves(dst, src, path="batch"):
    if torch.is_tensor(dst):
        if not torch.is_tensor(src):
            src = torch.as_tensor(src)
        if tuple(src.shape) != tuple(dst.shape):
            raise ValueError(f"{path}: shape {tuple(src.shape)} differs from the captured {tuple(dst.shape)} "
                             "(a captured step replays fixed shapes: pad or re-capture)")
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k, v in dst.items():
            _copy_leaves(v, src[k], f"{path}[{k!r}]")


class GraphedTrainStep(object):
    """The training step of train_itm.py:191-289 as ONE CUDA graph: both towers forward, the symmetric in-batch NLL (with its
    embedding all-gather under a process group), backward, gradient average, clip, AdamW, zero_grad - ~550 kernel launches
    and their Python dispatch replayed with one cudaGraphLaunch.  The eager step is launch-bound for a fifth of its time
    (measured: kernels cover 0.78-0.82 of the step); the replay is not.

        step = GraphedTrainStep(fwd_bwd, optimizer, batch, scheduler=sched)   # fwd_bwd(batch) -> loss, ends in loss.backward()
        for batch in loader: loss = step(batch)                                # same shapes as the captured batch

    What makes the captured launches follow the training state: dropout masks mix in a device word the graph bumps at
    the end of every replay (ldot_dropout_epoch), lr and the Adam bias corrections are read from device memory
    (FusedAdamW.device_hyper), inputs are copied into the static batch the graph reads.  `pos_ctx_indices` (a Python list
    in the reference's batches) is kept as a device tensor.  `warmup` eager steps run first (they are real steps): they
    lay out the optimiser's flat buffers and validate the batch (the checks a capture cannot perform); when the flat
    buffers do not exist yet one extra eager step is added (`layout_step`), so that the capture is preceded by an eager
    run of exactly the sequence it records - under a process group that is REQUIRED (NCCL connects lazily).
    Under a process group call release() before destroy_process_group() (see there).
    Do not keep losses of EARLIER eager steps alive with their grad_fn (store loss.detach()): autograd caches a
    parameter's AccumulateGrad node, with the stream it was created on, for as long as any graph references it, and ends
    every backward by joining those streams - inside a capture that is a dependency on uncaptured work."""

    def __init__(self, fwd_bwd, optimizer, batch, scheduler=None, warmup=2, layout_step=True):
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("GraphedTrainStep drives FusedAdamW (get_optimizer returns it)")
        lib = _lib.load()
        self.fwd_bwd, self.optimizer, self.scheduler = fwd_bwd, optimizer, scheduler
        dev = torch.device("cuda", torch.cuda.current_device())

        def static_of(t):
            return t.to(dev).clone() if t.is_cuda else t.to(dev, non_blocking=False)
        self.static = _map_leaves(dict(batch), static_of)
        if isinstance(self.static.get("pos_ctx_indices"), (list, tuple)):
            self.static["pos_ctx_indices"] = torch.tensor([int(v) for v in self.static["pos_ctx_indices"]],
                                                          dtype=torch.int64, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        optimizer.device_hyper(True)
        self.steps_taken = 0
        self.loss = None
        self.warmup_losses = []   # losses of the eager warm-up steps (device tensors)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            # The step that lays the optimiser's flat buffers out is not the steady-state step (no gradient sinks yet, no
            # overlapped gradient average): one more eager step runs the exact launch / collective sequence the capture
            # will record - NCCL establishes connections lazily at the first use of a collective, which must not happen
            # under capture.
            n_warm = max(1, int(warmup)) + (1 if (layout_step and optimizer._flat is None) else 0)
            for _ in range(n_warm):
                self.warmup_losses.append(self._eager().detach().clone())
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        _lib.check(lib.ldot_dropout_epoch(_lib.ptr(self.epoch)))
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self.fwd_bwd(self.static)
                self.optimizer.step()
                self.optimizer.zero_grad()
                self.epoch.add_(1)
        finally:
            _lib.check(lib.ldot_dropout_epoch(None))

    def release(self):
        """Destroy the captured graph and its memory pool.  REQUIRED before torch.distributed.destroy_process_group() when the
        step was captured under a process group: NCCL cannot tear a communicator down while a graph that captured its
        collectives still exists - destroy_process_group() then blocks forever (scripts/probes/graph_two_ranks.py)."""
        import gc
        self.graph = None
        self.loss = None
        self.static = None
        gc.collect()
        torch.cuda.synchronize()

    def _eager(self):
        loss = self.fwd_bwd(self.static)
        self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        self.optimizer.zero_grad()
        self.steps_taken += 1
        return loss

    def __call__(self, batch=None):
        """One replayed step on `batch` (None: the batch already in the static buffers).  -> the loss (a static device
        tensor, overwritten by the next replay)."""
        if batch is not None:
            _copy_leaves(self.static, batch)
        self.optimizer.begin_step()
        self.graph.replay()
        if self.scheduler is not None:
            self.scheduler.step()
        self.steps_taken += 1
        return self.loss
