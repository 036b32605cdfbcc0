"""Host-side mirror of dvl/options.py: the command-line / config-JSON surface of eval_itm.py and train_itm.py
(SURVEY.md Appendix C).  Same flag names, types, defaults and choices; same `--config file.json` semantics (JSON keys
become attributes unless the flag was given on the command line - dvl/options.py:96-109); same helpers.

  default_params / add_itm_params / add_logging_params / add_kd_params   dvl/options.py:15-93
  parse_with_config, map_db_dirs, print_args, set_seed, setup_args_gpu   dvl/options.py:96-183

The flag tables below are data (name, default, type | 'flag', choices); tests/golden/options_surface.json holds what
the reference's own parser produces for an empty command line and for its shipped config files, and
tests/test_host_logic.py compares this module against it.
"""
import argparse
import json
import logging
import os
import random
import socket
import sys

import numpy as np
import torch

logger = logging.getLogger()

FLAG = 'flag'   # action='store_true'

DEFAULT_PARAMS = [
    ('txt_model_type', 'bert-base', str), ('txt_model_config', 'bert-base', str), ('txt_checkpoint', None, str),
    ('img_model_type', 'uniter-base', str), ('img_model_config', './config/img_base.json', str),
    ('img_checkpoint', None, str), ('biencoder_checkpoint', None, str), ('seperate_caption_encoder', False, FLAG),
    ('train_batch_size', 80, int), ('valid_batch_size', 80, int), ('gradient_accumulation_steps', 1, int),
    ('learning_rate', 1e-5, float), ('max_grad_norm', 2.0, float), ('warmup_steps', 500, int), ('valid_steps', 500, int),
    ('num_train_steps', 5000, int), ('num_train_epochs', 0, int),
    ('fp16', False, FLAG), ('seed', 42, int), ('output_dir', './', str), ('max_txt_len', 64, int),
    ('local_rank', -1, int), ('config', None, str), ('itm_global_file', None, str), ('no_cuda', False, FLAG),
    ('n_workers', 2, int), ('pin_mem', False, FLAG), ('hnsw_index', False, FLAG), ('fp16_opt_level', 'O1', str),
    ('img_meta', None, str),
]

ITM_PARAMS = [
    ('conf_th', 0.2, float), ('caption_score_weight', 0.0, float), ('negative_size', 10, int),
    ('num_hard_negatives', 0, int), ('sample_init_hard_negatives', False, FLAG),
    ('hard_negatives_sampling', 'none', str, ['none', 'random', 'top', 'top-random', '10-20', '20-30']),
    ('max_bb', 100, int), ('min_bb', 10, int), ('num_bb', 36, int),
    ('train_txt_dbs', None, str), ('train_img_dbs', None, str),
    ('txt_db_mapping', None, str), ('img_db_mapping', None, str), ('pretrain_mapping', None, str),
    ('val_txt_db', None, str), ('val_img_db', None, str), ('test_txt_db', None, str), ('test_img_db', None, str),
    ('steps_per_hard_neg', -1, int), ('inf_minibatch_size', 400, int), ('project_dim', 0, int), ('cls_concat', '', str),
    ('fix_txt_encoder', False, FLAG), ('fix_img_encoder', False, FLAG), ('compressed_db', False, FLAG),
    ('retrieval_mode', 'both', str, ['img_only', 'txt_only', 'both']),
]

LOGGING_PARAMS = [('log_result_step', 4, int), ('project_name', 'itm', str), ('expr_name_prefix', '', str),
                  ('save_all_epochs', False, FLAG)]

KD_PARAMS = [('teacher_checkpoint', None, str), ('T', 1.0, float), ('kd_loss_weight', 1.0, float)]


def _add(parser, table):
    for entry in table:
        name, default, kind = entry[:3]
        if kind == FLAG:
            parser.add_argument('--' + name, action='store_true', help="")
        elif len(entry) > 3:
            parser.add_argument('--' + name, default=default, type=kind, choices=entry[3], help="")
        else:
            parser.add_argument('--' + name, default=default, type=kind, help="")


def default_params(parser: argparse.ArgumentParser):
    _add(parser, DEFAULT_PARAMS)


def add_itm_params(parser: argparse.ArgumentParser):
    _add(parser, ITM_PARAMS)


def add_logging_params(parser: argparse.ArgumentParser):
    _add(parser, LOGGING_PARAMS)


def add_kd_params(parser: argparse.ArgumentParser):
    _add(parser, KD_PARAMS)


def parse_with_config(parser, cmds=None):
    """dvl/options.py:96-109.  Keys of the --config JSON are set on the namespace unless the same flag appears on the
    process command line (the reference looks at sys.argv even when `cmds` is given; kept)."""
    args = parser.parse_args() if cmds is None else parser.parse_args(cmds)
    if args.config is not None:
        with open(args.config) as f:
            config_args = json.load(f)
        override_keys = {arg[2:].split('=')[0] for arg in sys.argv[1:] if arg.startswith('--')}
        for k, v in config_args.items():
            if k not in override_keys:
                setattr(args, k, v)
    return args


def map_db_dirs(args):
    """dvl/options.py:112-133: rewrite the '/pretrain', '/db', '/img' path prefixes through the *_mapping options."""
    rules = (('/pretrain', args.pretrain_mapping, 'pretrain'), ('/db', args.txt_db_mapping, 'db'),
             ('/img', args.img_db_mapping, 'img'))
    for k, v in list(vars(args).items()):
        if not isinstance(v, str):
            continue
        for prefix, target, tag in rules:
            v = vars(args)[k]
            if v.startswith(prefix) and target:
                print(tag, k, v)
                vars(args)[k] = v.replace(prefix, target)
    if args.img_db_mapping:
        args.train_img_dbs[:] = [p.replace('/img', args.img_db_mapping) for p in args.train_img_dbs]
    if args.txt_db_mapping:
        args.train_txt_dbs[:] = [p.replace('/db', args.txt_db_mapping) for p in args.train_txt_dbs]


def print_args(args):
    logger.info(" **************** CONFIGURATION **************** ")
    for key, val in sorted(vars(args).items()):
        logger.info("%s -->   %s", f"{key:<30}", val)
    logger.info(" **************** END CONFIGURATION **************** ")


def set_seed(args):
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if getattr(args, 'n_gpu', 0) > 0:
        torch.cuda.manual_seed_all(args.seed)


def setup_args_gpu(args):
    """dvl/options.py:151-183: device, n_gpu, distributed_world_size.  local_rank == -1 (or --no_cuda): single process;
    otherwise one process per GPU - the NCCL process group is initialised here (torchrun provides the rendezvous)."""
    if args.local_rank == -1 or args.no_cuda:
        device = torch.device("cuda" if torch.cuda.is_available() and not args.no_cuda else "cpu")
        args.n_gpu = torch.cuda.device_count()
    else:
        torch.cuda.set_device(args.local_rank)
        device = torch.device("cuda", args.local_rank)
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group(backend="nccl", device_id=device)
        args.n_gpu = 1
    args.device = device
    ws = os.environ.get('WORLD_SIZE')
    args.distributed_world_size = int(ws) if ws else 1
    logger.info('Initialized host %s as d.rank %d on device=%s, n_gpu=%d, world size=%d', socket.gethostname(),
                args.local_rank, device, args.n_gpu, args.distributed_world_size)
    logger.info("16-bits training: %s ", args.fp16)
