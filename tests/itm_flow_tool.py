"""TEST TOOLING for the drop-in boundary (SURVEY 8 b): everything needed to drive the eval_itm.py / train_itm.py flow over
a synthetic database directory, three ways -

  make_workspace(...)     database directories (synth.make_itm_db), tower config JSONs, a CheckpointState-style checkpoint
                          from seeded weights, and the eval / fine-tune config JSON in the shape of the reference's shipped
                          config/flickr30k_{eval,ft}_config.json
  eval_flow / train_flow  the call sequence of eval_itm.py:28-142 / train_itm.py:47-289 made through the namesake packages
                          (dvl, uniter_model, horovod) - what runs on the GPU box, where /root/reference does not exist
  python tests/itm_flow_tool.py [--cpu-doubles] SCRIPT ARGS...
                          run an UNMODIFIED reference script (lightningdot_b200.run_script); with --cpu-doubles the CUDA
                          pieces are replaced by the CPU oracle first, so that the host-side surface the script touches
                          (imports, flags, checkpoint loading, data layer, loaders, eval loop, optimiser / schedule calls)
                          is exercised in a container without a GPU.  The doubles live HERE (tests/), never in the product.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------------- workspace
def make_workspace(root, n_img=40, caps_per_img=5, layers=2, seed_db=3, seed_txt=301, seed_img=302, txt2img=None,
                   train=False, fp16=False, batch_size=16, n_workers=0, extra=None, seq_len=32, num_bb=36):
    """-> dict(config=path, checkpoint=path, txt_db=..., img_db=..., img2txt=...)."""
    from lightningdot_b200 import synth
    os.makedirs(root, exist_ok=True)
    txt_dir, img_dir = synth.make_itm_db(root, n_img, caps_per_img, seq_len=seq_len, num_bb=num_bb, seed=seed_db,
                                         txt2img=txt2img, name="val")
    tower_cfg = dict(vocab_size=synth.VOCAB, hidden_size=768, num_hidden_layers=layers, num_attention_heads=12,
                     intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                     attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                     initializer_range=0.02)
    cfg_txt, cfg_img = os.path.join(root, "txt_base.json"), os.path.join(root, "img_base.json")
    for p in (cfg_txt, cfg_img):
        with open(p, "w") as f:
            json.dump(tower_cfg, f)
    sd = {}
    for prefix, kind, seed in (("txt_model.", "txt", seed_txt), ("img_model.", "img", seed_img)):
        for k, v in synth.random_tower_state(kind, seed=seed, perturb=True, layers=layers).items():
            sd[prefix + k] = v
    ckpt_dir = os.path.join(root, "ckpt_run")
    os.makedirs(ckpt_dir, exist_ok=True)
    ckpt = os.path.join(ckpt_dir, "biencoder.last.pt")
    torch.save({"model_dict": sd, "optimizer_dict": None, "scheduler_dict": None, "offset": 0, "epoch": 0,
                "encoder_params": None}, ckpt)
    cfg = {"txt_model_type": "bert-base", "txt_model_config": cfg_txt, "img_model_type": "uniter-base",
           "img_model_config": cfg_img, "seed": 42, "output_dir": os.path.join(root, "out"), "max_txt_len": 60,
           "conf_th": 0.2, "max_bb": 100, "min_bb": 10, "num_bb": 36, "project_dim": 768,
           "val_txt_db": txt_dir, "val_img_db": img_dir, "project_name": "itm-synthetic", "n_workers": n_workers,
           "valid_batch_size": batch_size, "fp16": fp16}
    if train:
        cfg.update({"train_txt_dbs": [txt_dir], "train_img_dbs": [img_dir], "train_batch_size": batch_size,
                    "gradient_accumulation_steps": 1, "learning_rate": 2e-5, "num_train_epochs": 1,
                    "num_hard_negatives": 0, "hard_negatives_sampling": "none", "biencoder_checkpoint": ckpt,
                    "log_result_step": 2})
    cfg.update(extra or {})
    path = os.path.join(root, "train_config.json" if train else "eval_config.json")
    with open(path, "w") as f:
        json.dump(cfg, f, indent=1)
    with open(os.path.join(txt_dir, "img2txts.json")) as f:
        img2txt = json.load(f)
    return dict(config=path, checkpoint=ckpt, txt_db=txt_dir, img_db=img_dir, img2txt=img2txt)


# ----------------------------------------------------------------------------------------------------- the flows
def _parse(config, extra_argv=()):
    import argparse
    from dvl.options import add_itm_params, add_kd_params, add_logging_params, default_params, parse_with_config
    parser = argparse.ArgumentParser()
    default_params(parser)
    add_itm_params(parser)
    add_logging_params(parser)
    add_kd_params(parser)
    return parse_with_config(parser, ['--config', config] + list(extra_argv))


def _common_setup(args):
    from horovod import torch as hvd
    from transformers.tokenization_bert import BertTokenizer
    hvd.init()
    torch.cuda.set_device(hvd.local_rank())
    args.device = torch.device("cuda", hvd.local_rank())
    args.local_rank, args.n_gpu = hvd.rank(), hvd.size()
    args.tokenizer = BertTokenizer.from_pretrained(args.txt_model_config)
    args.vector_size = args.project_dim if args.project_dim > 0 else 768
    args.img_meta = None
    return args


def eval_flow(config, checkpoint, fp16=None, num_tops=100):
    """eval_itm.py:28-142 through the namesake packages -> dict(loss, acc, recall_txt, recall_img, rank_txt, rank_img,
    n_indexed).  (recall_txt here is the text -> image direction, as eval_model_on_dataloader names it.)"""
    from dvl.data.itm import itm_fast_collate
    from dvl.models.bi_encoder import BiEncoder, setup_for_distributed_mode
    from dvl.trainer import build_dataloader, eval_model_on_dataloader, load_dataset
    from uniter_model.data import ImageLmdbGroup
    args = _common_setup(_parse(config, ['--biencoder_checkpoint', checkpoint]))
    args.inf_minibatch_size = 400
    if fp16 is not None:
        args.fp16 = fp16
    bi_encoder = BiEncoder(args, args.fix_img_encoder, args.fix_txt_encoder, project_dim=args.project_dim)
    bi_encoder.load_state_dict(torch.load(args.biencoder_checkpoint, map_location='cpu')['model_dict'])
    for name in ("img_model", "txt_model"):
        m = getattr(bi_encoder, name)
        m.to(args.device)
        m, _ = setup_for_distributed_mode(m, None, args.device, args.n_gpu, -1, args.fp16, args.fp16_opt_level)
        m.eval()
    all_img_dbs = ImageLmdbGroup(args.conf_th, args.max_bb, args.min_bb, args.num_bb, args.compressed_db)
    dataset = load_dataset(all_img_dbs, args.val_txt_db, args.val_img_db, args, is_train=False)
    dataset.new_epoch()
    dataloader = build_dataloader(dataset, itm_fast_collate, False, args)
    with open(os.path.join(args.val_txt_db, 'img2txts.json')) as f:
        img2txt = json.load(f)
    loss, acc, (ix_img, ix_txt), (recall_txt, recall_img), (rank_txt, rank_img) = eval_model_on_dataloader(
        bi_encoder, dataloader, args, img2txt=img2txt, num_tops=num_tops)
    return dict(loss=loss, acc=acc, recall_txt=recall_txt, recall_img=recall_img, rank_txt=rank_txt, rank_img=rank_img,
                n_indexed=len(ix_img.index_id_to_db_id), indexers=(ix_img, ix_txt), bi_encoder=bi_encoder, args=args)


def train_flow(config, steps=2):
    """train_itm.py:47-289 through the namesake packages, `steps` optimiser steps of the first epoch -> list of losses +
    the model / optimiser for inspection."""
    from dvl.data.itm import itm_fast_collate
    from dvl.models.bi_encoder import (BiEncoder, BiEncoderNllLoss, get_optimizer, get_schedule_linear,
                                       load_biencoder_checkpoint, setup_for_distributed_mode)
    from dvl.trainer import build_dataloader, load_dataset
    from dvl.utils import _calc_loss
    from uniter_model.data import ImageLmdbGroup
    args = _common_setup(_parse(config))
    args.fp16_opt_level = 'O2'
    bi_encoder = BiEncoder(args, args.fix_img_encoder, args.fix_txt_encoder, args.project_dim)
    load_biencoder_checkpoint(bi_encoder, args.biencoder_checkpoint)
    optimizer = get_optimizer(bi_encoder, args.learning_rate)
    bi_encoder, optimizer = setup_for_distributed_mode(bi_encoder, optimizer, args.device, args.n_gpu, -1, args.fp16,
                                                       args.fp16_opt_level)
    all_img_dbs = ImageLmdbGroup(args.conf_th, args.max_bb, args.min_bb, args.num_bb, args.compressed_db)
    train_dataset = load_dataset(all_img_dbs, args.train_txt_dbs, args.train_img_dbs, args, True)
    for dset in train_dataset.datasets:
        dset.new_epoch(None, None)
    torch.manual_seed(args.seed)
    train_dataloader = build_dataloader(train_dataset, itm_fast_collate, True, args)
    total_updates = (len(train_dataloader) // args.gradient_accumulation_steps) * args.num_train_epochs
    scheduler = get_schedule_linear(optimizer, int(0.1 * total_updates), total_updates)
    bi_encoder.train()
    losses = []
    for step, batch in enumerate(train_dataloader):
        txt_vector, img_vectors, caption_vectors = bi_encoder(batch)
        loss_function = BiEncoderNllLoss()
        l_txt, c_txt, s_txt = _calc_loss(args, loss_function, img_vectors, txt_vector, caption_vectors,
                                         batch['pos_ctx_indices'], batch['neg_ctx_indices'], None)
        l_img, c_img, s_img = _calc_loss(args, loss_function, txt_vector, img_vectors, caption_vectors,
                                         batch['pos_ctx_indices'], batch['neg_ctx_indices'], None)
        loss = 0.5 * l_txt + 0.5 * l_img
        losses.append(loss.item())
        if args.fp16:
            from apex import amp
            with amp.scale_loss(loss, optimizer) as scaled_loss:
                scaled_loss.backward()
            torch.nn.utils.clip_grad_norm_(amp.master_params(optimizer), args.max_grad_norm)
        else:
            loss.backward()
            torch.nn.utils.clip_grad_norm_(bi_encoder.parameters(), args.max_grad_norm)
        optimizer.step()
        scheduler.step()
        bi_encoder.zero_grad()
        if step + 1 >= steps:
            break
    return dict(losses=losses, bi_encoder=bi_encoder, optimizer=optimizer, args=args)


# ----------------------------------------------------------------------------------------------------- CPU doubles
def install_cpu_doubles():
    """Replace every CUDA-touching piece with a CPU stand-in built on the oracle (test infrastructure)."""
    import lightningdot_b200.bi_encoder as be
    import lightningdot_b200.loader as loader
    import lightningdot_b200.trainer as trainer
    from oracle import flatip, loss as oloss, towers as otowers

    torch.cuda.set_device = lambda *a, **k: None
    _to = torch.nn.Module.to

    def to_cpu(self, *args, **kwargs):
        args = tuple(torch.device("cpu") if (isinstance(a, torch.device) and a.type == "cuda") or a == "cuda" else a
                     for a in args)
        return _to(self, *args, **kwargs)
    torch.nn.Module.to = to_cpu

    class OracleEngine:
        aliased = True

        def __init__(self, module):
            self.m = module

        def _sd(self):
            return dict(self.m.state_dict(keep_vars=True))

        def encode_text(self, ids, mask, pos, want_seq=False):
            seq, pooled = otowers.text_tower(self._sd(), ids, mask, pos)
            return (seq if want_seq else None), pooled

        def encode_image(self, ids, mask, pos, feat, box, gather_index=None, want_seq=False):
            seq, pooled = otowers.image_tower(self._sd(), ids, mask, pos, feat, box, gather_index)
            return (seq if want_seq else None), pooled

    be._TowerBase.engine = lambda self: OracleEngine(self)

    def train_forward(self, kind, inputs):   # autograd through the oracle towers (dropout off: deterministic double)
        eng = OracleEngine(self)
        return (eng.encode_text(*inputs[:3]) if kind == "txt" else eng.encode_image(*inputs))[1]
    be._TowerBase._train_forward = train_forward

    class OracleLoss:
        def calc(self, q, ctx, cap, pos, hard=None, caption_score_weight=0.1, experiment=None, reduction='mean'):
            return oloss.nll(q, ctx, pos, cap, caption_score_weight, reduction)
    be.BiEncoderNllLoss = trainer.BiEncoderNllLoss = OracleLoss
    be.get_optimizer = lambda model, learning_rate=1e-5, adam_eps=1e-8, weight_decay=0.0: torch.optim.AdamW(
        [p for p in model.parameters() if p.requires_grad], lr=learning_rate, eps=adam_eps, weight_decay=weight_decay)
    _setup = be.setup_for_distributed_mode

    def setup(model, optimizer, device, *a, **k):
        return _setup(model, optimizer, torch.device("cpu"), *a, **k)
    be.setup_for_distributed_mode = setup

    class OracleIndexer:
        def __init__(self, vector_sz, **kw):
            self.inner = flatip.FlatIndexer(vector_sz)
            self.index_id_to_db_id = self.inner.index_id_to_db_id

        def index_matrix(self, ids, vectors):
            self.inner.index_data(list(zip(ids, vectors.detach().cpu().numpy())))

        def search_knn(self, q, k):
            return self.inner.search_knn(q.detach().cpu().numpy() if isinstance(q, torch.Tensor) else q, k)
    trainer.DenseFlatIndexer = OracleIndexer

    class PassThroughLoader:
        def __init__(self, inner):
            self.loader = inner

        def __iter__(self):
            return iter(self.loader)

        def __len__(self):
            return len(self.loader)
    loader.PrefetchLoader = trainer.PrefetchLoader = PassThroughLoader
    import uniter_model.data as ud
    ud.PrefetchLoader = PassThroughLoader


def main():
    argv = sys.argv[1:]
    if argv and argv[0] == "--cpu-doubles":
        argv = argv[1:]
        install_cpu_doubles()
    from lightningdot_b200 import run_script
    run_script.main(argv)


if __name__ == "__main__":
    main()
