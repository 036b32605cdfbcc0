"""Host-side mirror of dvl/models/bi_encoder.py for the retrieval hot path.

Same names, constructor arguments, attributes, return values and state-dict keys as the reference:

  dot_product_scores            bi_encoder.py:54-68
  BertEncoder / UniterEncoder   bi_encoder.py:76-128 / 131-196   (.bert, .encode_proj, .config, init_encoder, get_out_size)
  BiEncoder                     bi_encoder.py:199-290            (.txt_model, .img_model, forward(batch))
  get_optimizer                 bi_encoder.py:566-576
  setup_for_distributed_mode    bi_encoder.py:579-610            (fp16=True selects the fp16 kernels; no apex)
  BiEncoderNllLoss              bi_encoder.py:613-665
  get_schedule_linear           bi_encoder.py:668-680
  load_biencoder_checkpoint     bi_encoder.py:737-752

The modules hold ordinary fp32 nn.Parameters under the reference's names (so reference checkpoints load with
strict=True and optimisers / state_dict() / .to() behave), but the arithmetic is NOT torch: forward() hands the
parameters to a TowerEngine (lightningdot_b200/towers.py) that runs the hand-written sm_100a kernels through the C
ABI.  There is no eager fallback: on a machine without the CUDA extension or a B200 the forward raises.
Out of scope here: BiEncoderForPretraining, BiEncoderForVisualQuestionAnswering, and the
training backward.
"""
import json
import logging
import os
from collections import defaultdict
from typing import Tuple

import torch
from torch import Tensor as T
from torch import nn
from torch.optim.lr_scheduler import LambdaLR

from . import _lib
from .towers import TowerEngine
from .training import FusedAdamW, NllFunction, run_tower_training

logger = logging.getLogger()

IMG_DIM = 2048  # dvl/const.py:1


# ------------------------------------------------------------------------------------------------------- configs
class TowerConfig(object):
    """The slice of BertConfig / UniterConfig (uniter_model/model/model.py:23-99) the towers read."""

    def __init__(self, vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12,
                 output_hidden_states=False, **unused):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.output_hidden_states = output_hidden_states
        if hidden_act != "gelu":
            raise ValueError("only the erf-GELU activation of the shipped configs is implemented")

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls(**json.load(reader))

    def to_dict(self):
        return dict(self.__dict__)

    def __repr__(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True)


# hub ids the shipped configs name (config/flickr30k_eval_config.json:2); resolved offline
_BUILTIN_CONFIGS = {
    "bert-base-cased": dict(vocab_size=28996),
    "bert-base-uncased": dict(vocab_size=30522),
    "bert-base": dict(vocab_size=28996),
}


def resolve_config(cfg_name):
    if cfg_name is None or cfg_name == "":
        cfg_name = "bert-base-uncased"
    if isinstance(cfg_name, TowerConfig):
        return cfg_name
    if cfg_name in _BUILTIN_CONFIGS:
        return TowerConfig(**_BUILTIN_CONFIGS[cfg_name])
    if os.path.isfile(cfg_name):
        return TowerConfig.from_json_file(cfg_name)
    raise ValueError(f"unknown model config '{cfg_name}': expected a JSON file or one of {sorted(_BUILTIN_CONFIGS)}")


# ------------------------------------------------------------------------------------------- parameter containers
class GELU(nn.Module):
    """uniter_model/model/layer.py:47-50 (erf form).  Only a state-dict placeholder inside encode_proj."""

    def forward(self, input_):
        raise RuntimeError("encode_proj runs inside the CUDA tower engine")


def _linear(i, o):
    return nn.Linear(i, o)


class _SelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query, self.key, self.value = _linear(c.hidden_size, c.hidden_size), _linear(c.hidden_size, c.hidden_size), \
            _linear(c.hidden_size, c.hidden_size)


class _SelfOutput(nn.Module):
    def __init__(self, c, inner):
        super().__init__()
        self.dense = _linear(inner, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=1e-12)


class _Attention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = _SelfAttention(c)
        self.output = _SelfOutput(c, c.hidden_size)


class _Intermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = _linear(c.hidden_size, c.intermediate_size)


class _Layer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _Attention(c)
        self.intermediate = _Intermediate(c)
        self.output = _SelfOutput(c, c.intermediate_size)


class _Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class _TextEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=1e-12)


class _ImageEmbeddings(nn.Module):
    def __init__(self, c, img_dim):
        super().__init__()
        self.img_linear = _linear(img_dim, c.hidden_size)
        self.img_layer_norm = nn.LayerNorm(c.hidden_size, eps=1e-12)
        self.pos_layer_norm = nn.LayerNorm(c.hidden_size, eps=1e-12)
        self.pos_linear = _linear(7, c.hidden_size)
        self.mask_embedding = nn.Embedding(2, img_dim, padding_idx=0)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=1e-12)


class _Pooler(nn.Module):
    """Present in checkpoints, computed-and-discarded by the reference (bi_encoder.py:116-120,187): never run."""

    def __init__(self, c):
        super().__init__()
        self.dense = _linear(c.hidden_size, c.hidden_size)


class _BertBody(nn.Module):
    def __init__(self, c, img_dim=None):
        super().__init__()
        self.embeddings = _TextEmbeddings(c)
        if img_dim is not None:
            self.img_embeddings = _ImageEmbeddings(c, img_dim)
        self.encoder = _Encoder(c)
        self.pooler = _Pooler(c)


def _make_proj(hidden, project_dim):
    return nn.Sequential(nn.Linear(hidden, hidden * 2), GELU(), nn.LayerNorm(hidden * 2, eps=1e-12),
                         nn.Linear(hidden * 2, project_dim))


class _TowerBase(nn.Module):
    KIND = None

    def __init__(self, config, project_dim: int = 0):
        super().__init__()
        assert config.hidden_size > 0, 'Encoder hidden_size can\'t be zero'
        self.config = config
        self.bert = _BertBody(config, IMG_DIM if self.KIND == "img" else None)
        self.encode_proj = _make_proj(config.hidden_size, project_dim) if project_dim > 0 else None
        self.apply(self._init_weights)
        self.compute_dtype = torch.bfloat16   # setup_for_distributed_mode(fp16=True) switches to torch.float16
        self._engine = None
        self._engine_sig = None
        self._engine_gen = -1

    def _init_weights(self, module):
        # uniter_model/model/model.py:134-147
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def get_out_size(self):
        if self.encode_proj:
            return self.encode_proj[-1].out_features
        return self.config.hidden_size

    # -- engine management -----------------------------------------------------------------------------------------
    def _signature(self):
        ps = list(self.parameters())
        return (self.compute_dtype, ps[0].device, tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps[:4]))

    def engine(self) -> TowerEngine:
        """(Re)build the 16-bit inference copies when parameters, device or compute dtype changed."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.LdotError("the towers run only on a CUDA device (B200): move the model with .to('cuda'); "
                                 "there is no CPU fallback")
        sig = self._signature()
        gen = _lib.param_generation[0]
        # (an engine whose weights alias the optimiser's buffers sees optimiser steps without a reload)
        if self._engine is None or self._engine_sig != sig or (not self._engine.aliased and self._engine_gen != gen):
            c = self.config
            eng = TowerEngine(self.KIND, c.hidden_size, c.num_attention_heads, c.intermediate_size, c.num_hidden_layers,
                              dtype=self.compute_dtype)
            eng.load(self.state_dict(), dev, trainable={n for n, p in self.named_parameters() if p.requires_grad})
            self._engine, self._engine_sig, self._engine_gen = eng, sig, gen
        return self._engine

    def zero_grad(self, set_to_none: bool = True):
        """nn.Module.zero_grad, except that gradients living in a FusedAdamW flat buffer are cleared IN PLACE (one
        memset per parameter group) so that the next backward keeps accumulating straight into them."""
        from .training import zero_grads
        zero_grads(self, set_to_none)

    def _wants_grad(self):
        """True when this call must be recorded for backward (train_itm.py:191-258): grad mode on and trainable
        parameters.  The training path keeps activations and attaches a TowerFunction node (training.py)."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _train_forward(self, kind, inputs):
        return run_tower_training(self, self.engine(), kind, inputs)


class BertEncoder(_TowerBase):
    """Text tower (bi_encoder.py:76-128): BERT-base + 768 -> 1536 -> project_dim head on the [CLS] state."""
    KIND = "txt"

    @classmethod
    def init_encoder(cls, cfg_name: str, checkpoint_path: str, project_dim: int = 0, dropout: float = 0.1, **kwargs):
        cfg = resolve_config(cfg_name)
        if dropout != 0:
            cfg.attention_probs_dropout_prob = dropout
            cfg.hidden_dropout_prob = dropout
        model = cls(cfg, project_dim=project_dim)
        if checkpoint_path is not None and len(checkpoint_path) > 0 and checkpoint_path.lower() != 'none':
            state_dict = torch.load(checkpoint_path, map_location='cpu')
            missing, unexpected = model.load_state_dict(state_dict, strict=False)
            logger.info(f'txt encoder loaded from {checkpoint_path}: missing {missing}, unexpected {unexpected}')
        else:
            logger.info('no checkpoint, random initialization for txt encoder')
        return model

    def forward(self, input_ids, attention_mask, position_ids,
                img_feat=None, img_pos_feat=None, img_masks=None, gather_index=None, need_sequence=True):
        """-> (sequence_output, pooled_output, hidden_states=None) as bi_encoder.py:107-123.  need_sequence=False (what
        BiEncoder.forward passes unless output_all_encoded_layers is set) returns sequence_output=None and lets the
        engine evaluate the last layer for the [CLS] position only; pooled_output is unchanged."""
        if self._wants_grad():
            if need_sequence:
                raise NotImplementedError("gradients flow through pooled_output only: call with need_sequence=False "
                                          "(BiEncoder.forward does) or under torch.no_grad()")
            return None, self._train_forward("txt", (input_ids, attention_mask, position_ids)), None
        seq, pooled = self.engine().encode_text(input_ids, attention_mask, position_ids, want_seq=need_sequence)
        hidden_states = None
        return seq, pooled, hidden_states


class UniterEncoder(_TowerBase):
    """Image-region tower (bi_encoder.py:131-196): UNITER-base over [CLS] + R region features."""
    KIND = "img"

    @classmethod
    def init_encoder(cls, cfg_name: str, checkpoint_path: str, project_dim: int = 0, dropout: float = 0.1, **kwargs):
        cfg = resolve_config(cfg_name)
        model = cls(cfg, project_dim=project_dim)
        if checkpoint_path is not None and len(checkpoint_path) > 0 and checkpoint_path.lower() != 'none':
            logger.info(f'load from {checkpoint_path} for uniter encoder')
            state_dict = torch.load(checkpoint_path, map_location='cpu')
            # uniter_model/model/model.py:165-177: legacy gamma / beta names
            for key in list(state_dict.keys()):
                new_key = key.replace('gamma', 'weight') if 'gamma' in key else key.replace('beta', 'bias') if 'beta' in key else None
                if new_key:
                    state_dict[new_key] = state_dict.pop(key)
            missing, unexpected = model.load_state_dict(state_dict, strict=False)
            logger.info(f'img encoder loaded: missing {missing}, unexpected {unexpected}')
        else:
            logger.info('no checkpoint, random initialization for img encoder')
        return model

    def forward(self, input_ids, attention_mask, position_ids,
                img_feat, img_pos_feat, img_masks, gather_index=None, need_sequence=True) -> Tuple[T, ...]:
        """bi_encoder.py:163-191; need_sequence as in BertEncoder.forward."""
        if img_masks is not None:
            raise NotImplementedError("img_masks (masked-region modelling, pre-training only) is outside the retrieval path")
        if self._wants_grad():
            if need_sequence:
                raise NotImplementedError("gradients flow through pooled_output only: call with need_sequence=False "
                                          "(BiEncoder.forward does) or under torch.no_grad()")
            if img_feat is None:
                return None, self._train_forward("txt", (input_ids, attention_mask, position_ids)), None
            return None, self._train_forward("img", (input_ids, attention_mask, position_ids, img_feat, img_pos_feat,
                                                     gather_index)), None
        eng = self.engine()
        if img_feat is None:   # txt_model_type == 'uniter-base': text through the UNITER body
            seq, pooled = eng.encode_text(input_ids, attention_mask, position_ids, want_seq=need_sequence)
        else:
            seq, pooled = eng.encode_image(input_ids, attention_mask, position_ids, img_feat, img_pos_feat,
                                           gather_index, want_seq=need_sequence)
        return seq, pooled, None


class BiEncoder(nn.Module):
    """ Bi-Encoder model component. Encapsulates query/question and context/passage encoders (bi_encoder.py:199-290)."""

    def __init__(self, args, fix_img_encoder: bool = False, fix_txt_encoder: bool = False, project_dim: int = 0):
        super(BiEncoder, self).__init__()
        logger.info('*' * 100)
        logger.info('loading img model')
        if args.img_model_type == 'uniter-base':
            self.img_model = UniterEncoder.init_encoder(args.img_model_config, checkpoint_path=args.img_checkpoint,
                                                        project_dim=project_dim)
        else:
            raise ValueError(f'image encoder does not support other types ({args.img_model_type}) for now')

        logger.info('*' * 100)
        logger.info('loading txt model')
        if args.txt_model_type == 'bert-base':
            self.txt_model = BertEncoder.init_encoder(args.txt_model_config, checkpoint_path=args.txt_checkpoint,
                                                      project_dim=project_dim)
        elif args.txt_model_type == 'uniter-base':
            self.txt_model = UniterEncoder.init_encoder(args.txt_model_config, checkpoint_path=args.txt_checkpoint,
                                                        project_dim=project_dim)
        else:
            raise ValueError(f'txt encoder does not support other types ({args.txt_model_type}) for now')

        self.fix_img_encoder = fix_img_encoder
        self.fix_txt_encoder = fix_txt_encoder
        self.project_dim = project_dim
        if fix_txt_encoder:
            for param in self.txt_model.parameters():
                param.requires_grad = False
        if fix_img_encoder:
            for param in self.img_model.parameters():
                param.requires_grad = False

    def zero_grad(self, set_to_none: bool = True):
        """bi_encoder.zero_grad() of train_itm.py:282: flat-buffer gradients are cleared in place (see _TowerBase)."""
        from .training import zero_grads
        zero_grads(self, set_to_none)

    @staticmethod
    def get_representation(sub_model, input_ids, attention_mask, position_ids, img_feat, img_pos_feat, img_masks,
                           gather_index=None, fix_encoder=False, need_sequence=True):
        if fix_encoder:
            with torch.no_grad():
                sequence_output, pooled_output, hidden_states = sub_model(input_ids, attention_mask, position_ids,
                                                                          img_feat, img_pos_feat, img_masks,
                                                                          gather_index, need_sequence=need_sequence)
        else:
            sequence_output, pooled_output, hidden_states = sub_model(input_ids, attention_mask, position_ids,
                                                                      img_feat, img_pos_feat, img_masks,
                                                                      gather_index, need_sequence=need_sequence)
        return sequence_output, pooled_output, hidden_states

    def forward(self, batch, output_all_encoded_layers=False):
        # batch keys: imgs / txts / caps  (dvl/data/itm.py:203-288)
        batch = defaultdict(lambda: None, batch)
        # the pooled outputs are all this call returns unless output_all_encoded_layers: skip the unread rows
        seq = bool(output_all_encoded_layers)

        if 'txts' in batch:
            sb = batch['txts']
            txt_seq, txt_pooled, txt_hidden = self.get_representation(self.txt_model, sb['input_ids'],
                                                                      sb['attention_mask'], sb['position_ids'],
                                                                      sb['img_feat'], sb['img_pos_feat'],
                                                                      sb['img_masks'],
                                                                      sb['gather_index'], self.fix_txt_encoder, need_sequence=seq)
        else:
            txt_seq, txt_pooled = None, None

        if 'imgs' in batch:
            sb = batch['imgs']
            # (the reference passes self.fix_txt_encoder here too, bi_encoder.py:273 - kept)
            img_seq, img_pooled, img_hidden = self.get_representation(self.img_model, sb['input_ids'],
                                                                      sb['attention_mask'], sb['position_ids'],
                                                                      sb['img_feat'], sb['img_pos_feat'],
                                                                      sb['img_masks'],
                                                                      sb['gather_index'], self.fix_txt_encoder, need_sequence=seq)
        else:
            img_seq, img_pooled = None, None

        if 'caps' in batch and batch['caps']['input_ids'] is not None:
            sb = batch['caps']
            cap_seq, cap_pooled, cap_hidden = self.get_representation(self.txt_model, sb['input_ids'],
                                                                      sb['attention_mask'], sb['position_ids'],
                                                                      sb['img_feat'], sb['img_pos_feat'],
                                                                      sb['img_masks'],
                                                                      sb['gather_index'], self.fix_txt_encoder, need_sequence=seq)
        else:
            cap_seq, cap_pooled = None, None

        if output_all_encoded_layers:
            return txt_seq, img_seq, cap_seq
        else:
            return txt_pooled, img_pooled, cap_pooled


# ------------------------------------------------------------------------------------------------------ scoring
def _scores_f32(q: T, ctx: T) -> T:
    """Q . C^T with fp32-grade accuracy on the 16-bit tensor cores (hi/lo split folded into one tcgen05 GEMM)."""
    lib = _lib.load()
    q = q.detach().to(torch.float32).contiguous()
    ctx = ctx.detach().to(device=q.device, dtype=torch.float32).contiguous()
    if q.dim() != 2 or ctx.dim() != 2 or q.shape[1] != ctx.shape[1]:
        raise ValueError(f"expected [n1, D] and [n2, D], got {tuple(q.shape)} and {tuple(ctx.shape)}")
    if q.device.type != "cuda":
        raise _lib.LdotError("dot_product_scores runs only on a CUDA device (B200); there is no CPU fallback")
    (n1, d), n2 = q.shape, ctx.shape[0]
    if d % 8 != 0:
        raise ValueError("vector size must be a multiple of 8")
    stream = _lib.stream_ptr()
    qs = torch.empty((n1, 3 * d), dtype=torch.float16, device=q.device)
    cs = torch.empty((n2, 3 * d), dtype=torch.float16, device=q.device)
    _lib.check(lib.ldot_split16(_lib.ptr(q), n1, d, 0, _lib.ptr(qs), stream))
    _lib.check(lib.ldot_split16(_lib.ptr(ctx), n2, d, 1, _lib.ptr(cs), stream))
    n2p = (n2 + 3) // 4 * 4   # fp32 output rows must stay 16-byte aligned
    out = torch.empty((n1, n2p), dtype=torch.float32, device=q.device)
    _lib.check(lib.ldot_linear(_lib.ptr(qs), 3 * d, _lib.ptr(cs), 3 * d, None, None, 0, _lib.ptr(out), n2p, n1, n2, 3 * d,
                               _lib.COARSE_FP16, 0, 1, stream))
    return out[:, :n2]


def dot_product_scores(q_vectors: T, ctx_vectors: T, cosine=False) -> T:
    """calculates q->ctx scores for every row in ctx_vector (bi_encoder.py:54-68): n1 x D, n2 x D -> n1 x n2"""
    r = _scores_f32(q_vectors, ctx_vectors)
    if cosine:
        n1 = torch.norm(q_vectors.float(), dim=-1)
        n2 = torch.norm(ctx_vectors.float(), dim=-1)
        return r / torch.outer(n1, n2)
    return r


class BiEncoderNllLoss(object):

    def calc(self, q_vectors: T, ctx_vectors: T, caption_vectors: T, positive_idx_per_question: list,
             hard_negatice_idx_per_question: list = None, caption_score_weight: float = 0.1,
             experiment=None, reduction='mean'):
        """
        Computes nll loss for the given lists of question and ctx vectors (bi_encoder.py:615-656).
        :return: a tuple of loss value and amount of correct predictions per batch (and the score matrix)
        """
        if reduction not in ('mean', 'sum'):
            raise ValueError("reduction must be 'mean' or 'sum'")
        # (hard_negatice_idx_per_question is accepted and unused, exactly as in the reference)
        dev = q_vectors.device
        if torch.is_tensor(positive_idx_per_question):
            # (a device tensor, e.g. the static index buffer of a captured step: range-checked unless a CUDA graph is
            # being captured, where a host read of device data is impossible)
            pos = positive_idx_per_question.to(device=dev, dtype=torch.int64).contiguous()
            if not (pos.is_cuda and torch.cuda.is_current_stream_capturing()) and pos.numel():
                if int(pos.min()) < 0 or int(pos.max()) >= ctx_vectors.shape[0]:
                    raise IndexError("positive index out of range")
        else:
            # the reference passes a Python list (dvl/utils.py:160-169): checked on the host, no device round trip
            idx = [int(v) for v in positive_idx_per_question]
            if idx and (min(idx) < 0 or max(idx) >= ctx_vectors.shape[0]):
                raise IndexError("positive index out of range")
            pos = torch.as_tensor(idx, dtype=torch.int64, device=dev).contiguous()
        if pos.numel() != q_vectors.shape[0]:
            raise ValueError("one positive index per question expected")
        w = float(caption_score_weight) if caption_vectors is not None else 0.0
        # one autograd node: scores + NLL forward, tensor-core backward to the embeddings (training.py: NllFunction)
        loss, correct, scores, scores_img = NllFunction.apply(q_vectors, ctx_vectors, caption_vectors if w != 0 else None,
                                                              pos, w, 0 if reduction == 'mean' else 1, self.get_scores)
        if experiment is not None:
            experiment.log_metric('score_img_diag_mean', torch.diag(scores_img).mean().item())
            experiment.log_metric('score_diag_mean', torch.diag(scores).mean().item())
        return loss, correct, scores

    @staticmethod
    def get_scores(q_vector: T, ctx_vectors: T) -> T:
        f = BiEncoderNllLoss.get_similarity_function()
        return f(q_vector, ctx_vectors)

    @staticmethod
    def get_similarity_function():
        return dot_product_scores


# --------------------------------------------------------------------------------------------- optimiser helpers
def get_optimizer(model: nn.Module, learning_rate: float = 1e-5, adam_eps: float = 1e-8,
                  weight_decay: float = 0.0, ) -> torch.optim.Optimizer:
    """bi_encoder.py:566-576 (transformers.AdamW == decoupled weight decay == torch.optim.AdamW)."""
    no_decay = ['bias', 'LayerNorm.weight']
    optimizer_grouped_parameters = [
        {'params': [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)],
         'weight_decay': weight_decay},
        {'params': [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], 'weight_decay': 0.0}
    ]
    # same update rule, one kernel launch per group over flat fp32 master buffers (training.py: FusedAdamW).  The model may
    # still be on the host here (train_itm.py builds the optimiser before setup_for_distributed_mode moves it): the
    # flat buffers are laid out at the first step(), when the parameters are on the device.
    return FusedAdamW(optimizer_grouped_parameters, lr=learning_rate, eps=adam_eps)


def setup_for_distributed_mode(model: nn.Module, optimizer: torch.optim.Optimizer, device: object, n_gpu: int = 1,
                               local_rank: int = -1,
                               fp16: bool = False,
                               fp16_opt_level: str = "O1",
                               teacher_model=None) -> (nn.Module, torch.optim.Optimizer):
    """bi_encoder.py:579-610.  The reference wraps the model with apex amp when fp16 is set; here fp16=True selects
    fp16 activations / weights for the tower kernels (what amp O1/O2 run the GEMMs in) and fp16=False the default
    bf16.  No apex needed.  With local_rank != -1 and an initialised torch.distributed group (one process per GPU)
    the reference would wrap the model in DistributedDataParallel (:603-607); here the parameters are broadcast from
    rank 0 once and the optimiser averages its flat gradient buffers over the ranks in step() (FusedAdamW.distributed),
    so the model object and its state_dict() keys stay unwrapped."""
    import torch.distributed as dist
    model.to(device)
    if teacher_model is not None:
        teacher_model.to(device)
    dtype = torch.float16 if fp16 else torch.bfloat16
    for m in model.modules():
        if isinstance(m, _TowerBase):
            m.compute_dtype = dtype
    if optimizer is not None and hasattr(optimizer, "shadow_dtype") and optimizer.shadow_dtype != dtype:
        optimizer.shadow_dtype = dtype
        optimizer._flat = None if not optimizer._flat else optimizer._relayout()
    if local_rank != -1 and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0)
        _lib.param_generation[0] += 1
        if optimizer is not None:
            if not hasattr(optimizer, "sync_gradients"):
                raise TypeError("distributed training needs the optimiser get_optimizer() returns (FusedAdamW)")
            optimizer.distributed = True
    return model, optimizer


def get_schedule_linear(optimizer, warmup_steps, training_steps, last_epoch=-1):
    """ Create a schedule with a learning rate that decreases linearly after
    linearly increasing during a warmup period (bi_encoder.py:668-680).
    """

    def lr_lambda(current_step):
        if current_step < warmup_steps:
            return float(current_step) / float(max(1, warmup_steps))
        return max(
            0.0, float(training_steps - current_step) / float(max(1, training_steps - warmup_steps))
        )

    return LambdaLR(optimizer, lr_lambda, last_epoch)


def load_biencoder_checkpoint(bi_encoder, biencoder_checkpoint):
    """bi_encoder.py:737-752: fine-tune checkpoints carry 'model_dict'; pre-training checkpoints prefix keys with
    'bert.' and carry extra heads that are dropped."""
    if biencoder_checkpoint is not None and len(biencoder_checkpoint) > 0 and biencoder_checkpoint.lower() != 'none':
        logger.info(f'loading ckpt from {biencoder_checkpoint}')
        state_dict = torch.load(biencoder_checkpoint, map_location='cpu')
        try:
            bi_encoder.load_state_dict(state_dict['model_dict'])
        except KeyError:
            logger.info('loading from pre-trained model instead')
            for k in list(state_dict.keys()):
                if k.startswith('bert.'):
                    state_dict[k[5:]] = state_dict.pop(k)
                else:
                    state_dict.pop(k)
            bi_encoder.load_state_dict(state_dict, strict=True)
    else:
        logger.info('no checkpoint provided, pass')
