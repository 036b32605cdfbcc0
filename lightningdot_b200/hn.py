"""Host-side mirror of dvl/hn.py: hard-negative mining for train_itm.py.

  random_hard_neg           dvl/hn.py:17-27    random same-set negatives
  get_img_txt_mappings      dvl/hn.py:30-44    img <-> txt <-> dataset-folder maps from the img2txts.json files
  sampled_hard_negatives    dvl/hn.py:47-68    encode the training set, search both directions with
                                               num_tops = min(max(2 * num_hard_negatives + 10, 50), 1000), drop the
                                               positives, sample num_hard_negatives per query

The device work is eval_model_on_dataloader (trainer.py): both towers over the whole training set, two exact top-k
searches with k up to 1000 through the same fused score + top-k kernel as evaluation (ldot_flatip_search supports
k <= 1024).  The per-dataset evaluation loaders come from load_dataset / build_dataloader (trainer.py, data.py) exactly
as in the reference; `train_dataloaders=` lets a caller pass ready-made loaders instead.
"""
import collections
import json
import logging
import os
import random
from collections import ChainMap

from .trainer import build_dataloader, eval_model_on_dataloader, load_dataset

logger = logging.getLogger()


def random_hard_neg(fname2id, num_hard_negatives, id2set, set2id):
    """dvl/hn.py:17-27: for every key, `num_hard_negatives` ids drawn (with replacement) from the key's own dataset,
    redrawn until the key's own id is not among them (rejection sampling: meant for very small counts)."""
    out = {}
    for key, own_id in fname2id.items():
        pool = set2id[id2set[key]]
        draw = random.choices(pool, k=num_hard_negatives)
        while own_id in draw:
            draw = random.choices(pool, k=num_hard_negatives)
        out[key] = draw
    return out


def get_img_txt_mappings(train_txt_dbs):
    """dvl/hn.py:30-44 -> (img2txt, txt2img, img2set, txt2set, set2img, set2txt) from the img2txts.json of every text
    database folder.  When an image appears in several folders the FIRST folder wins (ChainMap lookup order), and the
    per-dataset lists follow that resolved assignment."""
    per_db = []
    for folder in train_txt_dbs:
        with open(os.path.join(folder, 'img2txts.json')) as f:
            per_db.append(json.load(f))
    img2txt, img2set = {}, {}
    for folder, mapping in reversed(list(zip(train_txt_dbs, per_db))):   # later folders first, earlier ones overwrite
        for img, txts in mapping.items():
            img2txt[img] = txts
            img2set[img] = folder
    # (dict(ChainMap(...)) iterates the maps from the last to the first: reproduce that key order)
    txt2img = {t: img for img, txts in img2txt.items() for t in txts}
    txt2set = {t: img2set[img] for t, img in txt2img.items()}
    set2img, set2txt = collections.defaultdict(list), collections.defaultdict(list)
    for img, folder in img2set.items():
        set2img[folder].append(img)
        set2txt[folder].extend(img2txt[img])
    return img2txt, txt2img, img2set, txt2set, set2img, set2txt


def num_hard_sampled(num_hard_negatives):
    """Candidates retrieved per query before the positives are removed (dvl/hn.py:55)."""
    return min(max(num_hard_negatives * 2 + 10, 50), 1000)


def filter_and_sample(hard_neg_img, hard_neg_txt, train_img2txt, train_txt2img, num_hard_negatives):
    """dvl/hn.py:59-65.  hard_neg_img: {txt id: ranked image ids} (text -> image search), hard_neg_txt: {image id: ranked
    text ids}.  The positive image is removed from each text's list (in place, order kept), the positive captions from
    each image's list (through a set, as the reference does), then num_hard_negatives are drawn without replacement."""
    for k, v in hard_neg_img.items():
        if train_txt2img[k] in v:
            v.remove(train_txt2img[k])
    hard_neg_txt = {k: list(set(v) - set(train_img2txt[k])) for k, v in hard_neg_txt.items()}
    txt_out = {k: random.sample(v, num_hard_negatives) for k, v in hard_neg_txt.items()}
    img_out = {k: random.sample(v, num_hard_negatives) for k, v in hard_neg_img.items()}
    return txt_out, img_out


def sampled_hard_negatives(all_img_dbs, args, collate_func, bi_encoder, train_img2txt, train_txt2img,
                           train_dataloaders=None):
    """dvl/hn.py:47-68.  `train_dataloaders` (optional): one dataloader per training dataset; by default they are built
    as the reference does - load_dataset(..., is_train=True), new_epoch() without negatives, build_dataloader(dset,
    collate_func, True, args, args.valid_batch_size)."""
    if train_dataloaders is None:
        train_dataloaders = []
        for dset in load_dataset(all_img_dbs, args.train_txt_dbs, args.train_img_dbs, args, True).datasets:
            dset.new_epoch()
            train_dataloaders.append(build_dataloader(dset, collate_func, True, args, args.valid_batch_size))
    hard_negs_txt_all, hard_negs_img_all = [], []
    for loader in train_dataloaders:
        logger.info(f'eval for train dataloader len (for hn) = {len(loader)}')
        k = num_hard_sampled(args.num_hard_negatives)
        _, _, _, _, (hard_neg_img, hard_neg_txt) = eval_model_on_dataloader(bi_encoder, loader, args, train_img2txt, k)
        txt_out, img_out = filter_and_sample(hard_neg_img, hard_neg_txt, train_img2txt, train_txt2img,
                                             args.num_hard_negatives)
        hard_negs_txt_all.append(txt_out)
        hard_negs_img_all.append(img_out)
    return dict(ChainMap(*hard_negs_txt_all)), dict(ChainMap(*hard_negs_img_all))
