"""Host-side mirror of the data layer eval_itm.py / train_itm.py / rerank.py read their batches from.

  key-value stores           uniter_model/data/data.py:44-126,137-174   DetectFeatLmdb / TxtLmdb (LMDB environments)
  TxtTokLmdb                 uniter_model/data/data.py:177-214          id2len.json, meta.json, txt2img.json, img2txts.json
  DetectFeatTxtTokDataset    uniter_model/data/data.py:227-251
  ImageLmdbGroup             uniter_model/data/data.py:319-336
  pad_tensors, get_gather_index, get_ids_and_lens   uniter_model/data/data.py:217-224,270-294
  ItmFastDataset             dvl/data/itm.py:31-131      one (caption, image [, hard negatives] [, captions]) sample
  itm_fast_collate           dvl/data/itm.py:203-288     the nested batch BiEncoder.forward consumes
  ItmValDataset              dvl/data/itm.py:291-369     (rerank.py's mini-batch view)

This is the INPUT CONTRACT of the hot path, not an accelerated part of it: plain Python / torch-CPU, producing exactly
the tensors the reference produces.  tests/golden/itm_dataset_*.json is minted by running the reference's own classes
over the same database directory (oracle/make_golden.py) and compared byte for byte.

Storage.  The reference keeps records in LMDB; values are lz4-framed msgpack (text) and .npz / msgpack-numpy blobs
(image features).  `lmdb` and `lz4` are not installable in every environment (this image has neither), so a database
directory may hold either

  data.mdb          an LMDB environment (opened with the `lmdb` module when it is importable), or
  records.ldkv      a flat append-only file: b"LDKV1\\n" then per record  u32 key length | u64 value length | key | value

with the SAME value encodings.  Readers pick whichever exists; lz4 frames are recognised by their magic number and need
the `lz4` module, un-framed msgpack values are read directly.  The writers below (the counterpart of the reference's
prepro.py:384-411 and scripts/convert_imgdir.py:29-115) produce the JSON side files with the same names
and content, so a directory written here is a valid text / image database for the reference given its wheels.
"""
import io
import itertools
import json
import mmap
import os
import struct
from collections import defaultdict

import numpy as np
import torch
from torch.utils.data import Dataset

from .utils import get_rank, get_world_size

try:  # optional wheels (absent in this image)
    import lmdb as _lmdb
except Exception:  # pragma: no cover
    _lmdb = None
try:
    from lz4 import frame as _lz4frame
except Exception:  # pragma: no cover
    _lz4frame = None
import msgpack

FLAT_NAME = "records.ldkv"
_FLAT_MAGIC = b"LDKV1\n"
_LZ4_MAGIC = b"\x04\x22\x4d\x18"
N_EXAMPLES_TEACHER = 10   # GLOBAL_VARIABLES.py:6


# ------------------------------------------------------------------------------------------------ key-value stores
class FlatKV(object):
    """records.ldkv reader / writer.  Reading maps the file and keeps {key: (offset, length)}; fork-safe (DataLoader
    workers share the read-only mapping)."""

    def __init__(self, db_dir, write=False):
        self.path = os.path.join(db_dir, FLAT_NAME)
        self.write = write
        self._index = {}
        if write:
            os.makedirs(db_dir, exist_ok=True)
            self._f = open(self.path, "wb")
            self._f.write(_FLAT_MAGIC)
            self._map = None
        else:
            self._f = open(self.path, "rb")
            self._map = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
            if self._map[:len(_FLAT_MAGIC)] != _FLAT_MAGIC:
                raise ValueError(f"{self.path}: not a records.ldkv file")
            pos, end = len(_FLAT_MAGIC), len(self._map)
            while pos < end:
                klen, vlen = struct.unpack_from("<IQ", self._map, pos)
                pos += 12
                key = bytes(self._map[pos:pos + klen])
                pos += klen
                self._index[key] = (pos, vlen)
                pos += vlen

    def get(self, key):
        hit = self._index.get(key)
        if hit is None:
            return None
        off, n = hit
        return self._map[off:off + n]

    def put(self, key, value):
        self._f.write(struct.pack("<IQ", len(key), len(value)))
        self._f.write(key)
        self._f.write(value)

    def keys(self):
        return list(self._index.keys())

    def close(self):
        if self._map is not None:
            self._map.close()
            self._map = None
        if self._f is not None:
            self._f.close()
            self._f = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LmdbKV(object):
    """The reference's storage (uniter_model/data/data.py:66-69,143-152): read-only LMDB environment, one long-lived
    read transaction with buffers=True."""

    def __init__(self, db_dir, write=False, readahead=False):
        if _lmdb is None:
            raise ImportError(f"{db_dir} is an LMDB environment and the `lmdb` module is not installed")
        self.write = write
        if write:
            self.env = _lmdb.open(db_dir, readonly=False, create=True, map_size=4 * 1024 ** 4)
            self.txn = self.env.begin(write=True)
        else:
            self.env = _lmdb.open(db_dir, readonly=True, create=False, readahead=readahead)
            self.txn = self.env.begin(buffers=True)

    def get(self, key):
        return self.txn.get(key)

    def put(self, key, value):
        self.txn.put(key, value)

    def keys(self):
        return [bytes(k) for k, _ in self.txn.cursor()]

    def close(self):
        if self.write:
            self.txn.commit()
        self.env.close()


def open_kv(db_dir, write=False, backend=None):
    """Open the record store of a database directory.  Reading: LMDB when data.mdb is there, else records.ldkv.
    Writing: `backend` in {'lmdb', 'flat'}, default lmdb when the module is importable."""
    if write:
        backend = backend or ("lmdb" if _lmdb is not None else "flat")
        return LmdbKV(db_dir, write=True) if backend == "lmdb" else FlatKV(db_dir, write=True)
    if os.path.exists(os.path.join(db_dir, "data.mdb")):
        # (the reference turns read-ahead off for multi-node runs only, data.py:36-41,68; it never helps random access)
        return LmdbKV(db_dir)
    if os.path.exists(os.path.join(db_dir, FLAT_NAME)):
        return FlatKV(db_dir)
    raise FileNotFoundError(f"{db_dir}: neither data.mdb (LMDB) nor {FLAT_NAME}")


# ---- value codecs
def _nd_encode(obj):
    """msgpack-numpy's wire form of an ndarray [external: msgpack_numpy.encode]."""
    if isinstance(obj, np.ndarray):
        return {"nd": True, "type": obj.dtype.str, "kind": "", "shape": list(obj.shape), "data": obj.tobytes()}
    if isinstance(obj, np.generic):
        return {"nd": False, "type": obj.dtype.str, "data": obj.tobytes()}
    return obj


def _nd_decode(obj):
    if isinstance(obj, dict) and "nd" in obj and "type" in obj and "data" in obj:
        if obj["nd"]:
            return np.frombuffer(obj["data"], dtype=np.dtype(obj["type"])).reshape(obj["shape"])
        return np.frombuffer(obj["data"], dtype=np.dtype(obj["type"]))[0]
    return obj


def pack_record(value, compress=True):
    """Text-database value: msgpack, lz4-framed when the module is there (data.py:166-167)."""
    raw = msgpack.dumps(value, use_bin_type=True, default=_nd_encode)
    return _lz4frame.compress(raw) if (compress and _lz4frame is not None) else raw


def unpack_record(blob):
    blob = bytes(blob)
    if blob[:4] == _LZ4_MAGIC:
        if _lz4frame is None:
            raise ImportError("this text database is lz4-compressed and the `lz4` module is not installed")
        blob = _lz4frame.decompress(blob)
    return msgpack.loads(blob, raw=False, object_hook=_nd_decode)


def pack_features(arrays, compress=True):
    """Image-database value: an .npz blob for the `*_compressed` databases (convert_imgdir.py:40-47), msgpack-numpy
    otherwise."""
    if compress:
        with io.BytesIO() as w:
            np.savez_compressed(w, **arrays)
            return w.getvalue()
    return msgpack.dumps(dict(arrays), use_bin_type=True, default=_nd_encode)


def unpack_features(blob, compress=True, fields=None):
    if compress:
        with io.BytesIO(bytes(blob)) as r:
            z = np.load(r, allow_pickle=True)
            return {k: z[k] for k in (fields or z.files)}
    d = msgpack.loads(bytes(blob), raw=False, object_hook=_nd_decode)
    return d if fields is None else {k: d[k] for k in fields}


def compute_num_bb(confs, conf_th, min_bb, max_bb):
    """data.py:30-33: boxes above the confidence threshold, clamped to [min_bb, max_bb]."""
    return int(min(max_bb, max(min_bb, int((np.asarray(confs) > conf_th).sum()))))


def _img_db_name(conf_th, max_bb, min_bb, num_bb, compress, have_nbb=True):
    if conf_th == -1:
        name = f"feat_numbb{num_bb}"
    elif have_nbb:
        name = f"feat_th{conf_th}_max{max_bb}_min{min_bb}"
    else:
        name = "all"
    return name + ("_compressed" if compress else "")


# ------------------------------------------------------------------------------------------------ image features
class DetectFeatLmdb(object):
    """Region-feature database of one image folder (data.py:44-126): `db[fname] -> (features [nbb, 2048] fp32,
    norm_bb [nbb, 6] fp32)`, `db.name2nbb[fname]`, `fname in db`, `db.get_dump(fname)`."""

    def __init__(self, img_dir, conf_th=0.2, max_bb=100, min_bb=10, num_bb=36, compress=True):
        self.img_dir = img_dir
        self.conf_th, self.max_bb, self.min_bb, self.num_bb = conf_th, max_bb, min_bb, num_bb
        self.compress = compress
        have_nbb = True
        if conf_th == -1:
            self.name2nbb = defaultdict(lambda: num_bb)
        else:
            nbb_file = os.path.join(img_dir, f"nbb_th{conf_th}_max{max_bb}_min{min_bb}.json")
            if os.path.exists(nbb_file):
                with open(nbb_file) as f:
                    self.name2nbb = json.load(f)
            else:
                self.name2nbb, have_nbb = None, False
        self.db_name = _img_db_name(conf_th, max_bb, min_bb, num_bb, compress, have_nbb)
        self.kv = open_kv(os.path.join(img_dir, self.db_name))
        if self.name2nbb is None:   # box counts not pre-computed: derive them from the stored confidences
            names = json.loads(bytes(self.kv.get(b"__keys__")).decode("utf-8"))
            self.name2nbb = {n: compute_num_bb(self._load(n, ("conf",))["conf"], conf_th, min_bb, max_bb) for n in names}

    def _load(self, file_name, fields=None):
        blob = self.kv.get(file_name.encode("utf-8"))
        if blob is None:
            raise KeyError(file_name)
        return unpack_features(blob, self.compress, fields)

    def get_dump(self, file_name):
        nbb = self.name2nbb[file_name]
        out = {}
        for k, arr in self._load(file_name).items():
            arr = np.asarray(arr)
            out[k] = (arr.astype(np.float32) if arr.dtype == np.float16 else arr)[:nbb, ...]
        return out

    def __getitem__(self, file_name):
        d = self._load(file_name, ("features", "norm_bb"))
        nbb = self.name2nbb[file_name]
        return (torch.tensor(np.asarray(d["features"])[:nbb, :]).float(),
                torch.tensor(np.asarray(d["norm_bb"])[:nbb, :]).float())

    def __contains__(self, file_name):
        return self.kv.get(file_name.encode("utf-8")) is not None


class ImageLmdbGroup(object):
    """data.py:319-336: image-folder path -> DetectFeatLmdb with one shared box-count policy.  (The reference never
    fills its cache, so every lookup re-opens the database; here an opened database is kept.)"""

    def __init__(self, conf_th, max_bb, min_bb, num_bb, compress):
        self.path2imgdb = {}
        self.conf_th, self.max_bb, self.min_bb, self.num_bb, self.compress = conf_th, max_bb, min_bb, num_bb, compress

    def __getitem__(self, path):
        db = self.path2imgdb.get(path)
        if db is None:
            db = DetectFeatLmdb(path, self.conf_th, self.max_bb, self.min_bb, self.num_bb, self.compress)
            self.path2imgdb[path] = db
        return db


# ------------------------------------------------------------------------------------------------ text records
class TxtLmdb(object):
    """data.py:137-174: `db[id] -> record dict` ({'input_ids': [...], 'img_fname': ..., ...})."""

    def __init__(self, db_dir, readonly=True):
        self.readonly = readonly
        self.kv = open_kv(db_dir, write=not readonly)

    def __getitem__(self, key):
        blob = self.kv.get(key.encode("utf-8"))
        if blob is None:
            raise KeyError(key)
        return unpack_record(blob)

    def __setitem__(self, key, value):
        if self.readonly:
            raise ValueError("readonly text DB")
        self.kv.put(key.encode("utf-8"), pack_record(value))

    def close(self):
        self.kv.close()


class TxtTokLmdb(object):
    """data.py:177-214: tokenised captions of one split.  `ids` are the caption ids no longer than max_txt_len (-1: all),
    strided over the ranks of the job exactly as the reference does (ids[rank::size])."""

    def __init__(self, db_dir, max_txt_len=60):
        with open(os.path.join(db_dir, "id2len.json")) as f:
            self.id2len = json.load(f)
        ids = [i for i, n in self.id2len.items() if max_txt_len == -1 or n <= max_txt_len]
        self.ids = ids[get_rank()::get_world_size()]
        self.db_dir = db_dir
        self.db = TxtLmdb(db_dir, readonly=True)
        with open(os.path.join(db_dir, "meta.json")) as f:
            meta = json.load(f)
        self.cls_, self.sep, self.mask, self.v_range = meta["CLS"], meta["SEP"], meta["MASK"], meta["v_range"]

    def __getitem__(self, id_):
        return self.db[id_]

    def combine_inputs(self, *inputs):
        """[CLS] seg_1 [SEP] seg_2 [SEP] ... as an int64 tensor."""
        out = [self.cls_]
        for seg in inputs:
            out.extend(seg)
            out.append(self.sep)
        return torch.tensor(out)

    def _side(self, name):
        with open(os.path.join(self.db_dir, name)) as f:
            return json.load(f)

    @property
    def txt2img(self):
        return self._side("txt2img.json")

    @property
    def img2txts(self):
        return self._side("img2txts.json")


def get_ids_and_lens(db):
    assert isinstance(db, TxtTokLmdb)
    return [db.id2len[i] for i in db.ids], list(db.ids)


class DetectFeatTxtTokDataset(Dataset):
    """data.py:227-251: caption i of the text database + region features by file name."""

    def __init__(self, txt_db, img_db):
        assert isinstance(txt_db, TxtTokLmdb)
        assert isinstance(img_db, DetectFeatLmdb)
        self.txt_db, self.img_db = txt_db, img_db
        txt_lens, self.ids = get_ids_and_lens(txt_db)
        txt2img = txt_db.txt2img
        self.lens = [tl + self.img_db.name2nbb[txt2img[i]] for tl, i in zip(txt_lens, self.ids)]

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        return self.txt_db[self.ids[i]]

    def _get_img_feat(self, fname):
        """-> (features [nbb, 2048], boxes [nbb, 7] = x1 y1 x2 y2 w h w*h, nbb)"""
        feat, bb = self.img_db[fname]
        return feat, torch.cat([bb, bb[:, 4:5] * bb[:, 5:]], dim=-1), feat.size(0)


def pad_tensors(tensors, lens=None, pad=0):
    """B x [T_i, D] -> [B, max T, D], rows past T_i filled with `pad` (data.py:270-283)."""
    lens = [t.size(0) for t in tensors] if lens is None else lens
    out = torch.full((len(tensors), max(lens), tensors[0].size(-1)), pad, dtype=tensors[0].dtype)
    for row, (t, n) in enumerate(zip(tensors, lens)):
        out[row, :n] = t
    return out


def get_gather_index(txt_lens, num_bbs, batch_size, max_len, out_size):
    """data.py:286-294: the identity gather the bi-encoder batches carry (one arange(out_size) row per image)."""
    return torch.arange(out_size, dtype=torch.long).unsqueeze(0).repeat(len(num_bbs), 1)


# ------------------------------------------------------------------------------------------------ ITM samples
_CLS_IMG = 101   # the image tower's single text token (dvl/data/itm.py:73)


class ItmFastDataset(DetectFeatTxtTokDataset):
    """dvl/data/itm.py:31-131.  Sample i = caption i, its image, optionally `num_hard_negatives` mined negatives for
    both, optionally the image's captions concatenated (img_meta + tokenizer).  new_epoch() must be called before
    iterating; it fixes the image of every caption and the negatives for this epoch."""

    def __init__(self, txt_db, img_db, num_hard_negatives=0, img_meta=None, tokenizer=None):
        assert isinstance(txt_db, TxtTokLmdb)
        assert isinstance(img_db, DetectFeatLmdb)
        self.txt_db, self.img_db = txt_db, img_db
        self.txt_lens, self.ids = get_ids_and_lens(txt_db)
        self.ids_2_idx = {id_: i for i, id_ in enumerate(self.ids)}
        self._img_of = [txt_db[id_]['img_fname'] for id_ in self.ids]
        self.all_imgs = list(set(self._img_of))
        self.num_hard_negatives = num_hard_negatives
        self.img_meta, self.tokenizer = img_meta, tokenizer
        self.train_imgs = self.neg_imgs = None

    def new_epoch(self, hard_negatives_img=None, hard_negatives_txt=None):
        k = self.num_hard_negatives
        mined = hard_negatives_img is not None and k > 0
        self.train_imgs, self.train_txts = list(self._img_of), list(self.ids)
        self.neg_imgs = [hard_negatives_img[t][:k] if mined else None for t in self.ids]
        self.neg_txts = [hard_negatives_txt[f][:k] if mined else None for f in self._img_of]
        self.lens = [tl + self.img_db.name2nbb[f] for tl, f in zip(self.txt_lens, self._img_of)]

    def _caption_ids(self, fname, like):
        """[CLS] cap_1 [SEP] cap_2 [SEP] ... over the image's captions (itm.py:116-118)."""
        tok = self.tokenizer
        pieces = [tok.encode(c, add_special_tokens=False) + [tok.sep_token_id]
                  for c in self.img_meta[fname]['caption_multiple']]
        return torch.tensor([tok.cls_token_id] + list(itertools.chain.from_iterable(pieces)), dtype=like.dtype)

    def _text_ids(self, i):
        return self.txt_db.combine_inputs(DetectFeatTxtTokDataset.__getitem__(self, i)['input_ids'])

    def __getitem__(self, i):
        img_fname, neg_img_names, neg_txt_names = self.train_imgs[i], self.neg_imgs[i], self.neg_txts[i]
        img_feat, img_pos_feat, num_bb = self._get_img_feat(img_fname)
        input_ids = self._text_ids(i)
        neg_imgs = neg_txts = None
        if neg_img_names is not None:
            neg_imgs = {k: [] for k in ('img_input_ids', 'img_feat', 'img_pos_feat', 'num_bb', 'attn_masks_img',
                                        'caption_ids', 'attn_masks_captions')}
            for name in neg_img_names:
                f, p, n = self._get_img_feat(name)
                neg_imgs['img_input_ids'].append(torch.tensor([_CLS_IMG], dtype=torch.long))
                neg_imgs['img_feat'].append(f)
                neg_imgs['img_pos_feat'].append(p)
                neg_imgs['num_bb'].append(n)
                neg_imgs['attn_masks_img'].append(torch.ones(n + 1, dtype=torch.long))
                if self.img_meta is not None:
                    c = self._caption_ids(name, input_ids)
                    neg_imgs['caption_ids'].append(c)
                    neg_imgs['attn_masks_captions'].append(torch.ones(len(c), dtype=torch.long))
            neg_txts = {'input_ids': [], 'position_ids': [], 'attention_mask': []}
            for tid in neg_txt_names:
                t = self._text_ids(self.ids_2_idx[tid])
                neg_txts['input_ids'].append(t)
                neg_txts['attention_mask'].append(torch.ones(len(t), dtype=torch.long))
        caption_ids = attn_masks_captions = None
        if self.img_meta is not None:
            caption_ids = self._caption_ids(img_fname, input_ids)
            attn_masks_captions = torch.ones(len(caption_ids), dtype=torch.long)
        return (input_ids, img_feat, img_pos_feat, torch.tensor([_CLS_IMG], dtype=torch.long),
                torch.ones(len(input_ids), dtype=torch.long), torch.ones(num_bb + 1, dtype=torch.long),
                self.ids[i], img_fname, neg_imgs, neg_txts, caption_ids, attn_masks_captions)


def _pad_rows(seqs, value=0):
    """List of 1-D int64 tensors -> [len, max] padded with `value`."""
    return torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True, padding_value=value)


def _chain(dicts, key):
    return list(itertools.chain.from_iterable(d[key] for d in dicts))


def itm_fast_collate(inputs):
    """dvl/data/itm.py:203-288.  `inputs`: ItmFastDataset samples.  Positives first, then every sample's hard negatives
    (images after the batch's images, texts after the batch's texts); 'pos_ctx_indices' = arange(batch),
    'neg_ctx_indices' = the rest of the image rows."""
    cols = list(zip(*inputs))
    (input_ids, img_feats, img_pos_feats, img_input_ids, masks_txt, masks_img, idx, img_fname, neg_imgs, neg_txts,
     caption_ids, masks_cap) = [list(c) for c in cols]
    bs = len(input_ids)
    if None not in neg_imgs:
        nbb_neg = _chain(neg_imgs, 'num_bb')
        img_feats = img_feats + _chain(neg_imgs, 'img_feat')
        img_pos_feats = img_pos_feats + _chain(neg_imgs, 'img_pos_feat')
        img_input_ids = img_input_ids + _chain(neg_imgs, 'img_input_ids')
        masks_img = masks_img + _chain(neg_imgs, 'attn_masks_img')
        caption_ids = caption_ids + _chain(neg_imgs, 'caption_ids')
        masks_cap = masks_cap + _chain(neg_imgs, 'attn_masks_captions')
        input_ids = input_ids + _chain(neg_txts, 'input_ids')
        masks_txt = masks_txt + _chain(neg_txts, 'attention_mask')
    else:
        nbb_neg = []
    has_caps = caption_ids[0] is not None
    num_bbs = [f.size(0) for f in img_feats[:bs]] + nbb_neg
    txt = _pad_rows(input_ids)
    caps = _pad_rows(caption_ids) if has_caps else None
    mask_img = _pad_rows(masks_img)
    none4 = {'img_feat': None, 'img_pos_feat': None, 'img_masks': None, 'gather_index': None}

    def positions(t):
        return None if t is None else torch.arange(t.size(1), dtype=torch.long).unsqueeze(0)

    img_ids = _pad_rows(img_input_ids)
    return {
        'txts': {'input_ids': txt, 'position_ids': positions(txt), 'attention_mask': _pad_rows(masks_txt), **none4},
        'imgs': {'input_ids': img_ids, 'position_ids': positions(img_ids), 'attention_mask': mask_img,
                 'img_feat': pad_tensors(img_feats, num_bbs), 'img_pos_feat': pad_tensors(img_pos_feats, num_bbs),
                 'img_masks': None, 'gather_index': get_gather_index([1] * bs, num_bbs, bs, 1, mask_img.size(1))},
        'caps': {'input_ids': caps, 'position_ids': positions(caps),
                 'attention_mask': _pad_rows(masks_cap) if has_caps and masks_cap[0] is not None else None, **none4},
        'sample_size': bs,
        'pos_ctx_indices': list(range(bs)),
        'neg_ctx_indices': list(range(bs, len(num_bbs))),
        'txt_index': idx,
        'img_fname': img_fname,
    }


def itm_fast_collate_kd(inputs):
    """dvl/data/itm.py:134-200 builds the extra cross-encoder teacher inputs of the knowledge-distillation mode
    (train_itm.py:85-95,224-241).  The teacher (UniterForImageTextRetrieval) is outside the bi-encoder retrieval path
    (SURVEY 2.1), so is this collate; the name exists because train_itm.py imports it."""
    raise NotImplementedError("knowledge distillation from a UNITER cross-encoder teacher (--teacher_checkpoint) is outside "
                              "the retrieval hot path this package implements")


class ItmValDataset(DetectFeatTxtTokDataset):
    """dvl/data/itm.py:291-369: item i = caption i against a mini-batch of images (its own first, then the next
    mini_batch_size - 1 images in database order, wrapping around) - the cross-encoder evaluation view."""

    def __init__(self, db_dir, img_dir, mini_batch_size=400):
        super().__init__(db_dir, img_dir)
        del self.lens
        self.txt2img = self.txt_db.txt2img
        self.img2txts = self.txt_db.img2txts
        self.all_img_ids = list(self.img2txts.keys())
        assert len(self.img2txts) >= mini_batch_size > 0
        self.bs = mini_batch_size

    def _get_batch_ids(self, i):
        gt = self.txt2img[self.ids[i]]
        at, n = self.all_img_ids.index(gt), len(self.all_img_ids)
        return gt, [self.all_img_ids[(at + 1 + j) % n] for j in range(self.bs - 1)]   # (wraps around the end)

    def __getitem__(self, i):
        gt, negs = self._get_batch_ids(i)
        return self.get_batch(i, [gt] + negs)

    def get_batch(self, i, img_ids):
        ids = self.txt_db.combine_inputs(DetectFeatTxtTokDataset.__getitem__(self, i)['input_ids'])
        n = len(img_ids)
        input_ids = ids.unsqueeze(0).expand(n, -1).clone()
        feats, boxes, num_bbs = zip(*[self._get_img_feat(f) for f in img_ids])
        num_bbs = list(num_bbs)
        mask_img = torch.zeros(n, max(num_bbs), dtype=torch.long)
        for r, nbb in enumerate(num_bbs):
            mask_img[r, :nbb] = 1
        return {'input_ids': input_ids,
                'position_ids': torch.arange(input_ids.size(1), dtype=torch.long).unsqueeze(0),
                'img_feat': pad_tensors(list(feats), num_bbs), 'img_pos_feat': pad_tensors(list(boxes), num_bbs),
                'attn_masks_text': torch.ones(n, input_ids.size(1), dtype=torch.long),
                'attn_masks_img': mask_img, 'gather_index': None}


# ------------------------------------------------------------------------------------------------ writers
def write_txt_db(db_dir, records, cls_id=101, sep_id=102, mask_id=103, v_range=(106, 28996), backend=None):
    """Write a tokenised-caption database directory (prepro.py:384-411 [layout]): records = {caption id: dict with
    'input_ids' (no special tokens) and 'img_fname' (+ anything else)} in insertion order.  Produces the record store,
    id2len.json, txt2img.json, img2txts.json and meta.json."""
    os.makedirs(db_dir, exist_ok=True)
    kv = open_kv(db_dir, write=True, backend=backend)
    id2len, txt2img, img2txts = {}, {}, {}
    for key, rec in records.items():
        kv.put(key.encode("utf-8"), pack_record(rec))
        id2len[key] = len(rec['input_ids'])
        txt2img[key] = rec['img_fname']
        img2txts.setdefault(rec['img_fname'], []).append(key)
    kv.close()
    for name, obj in (("id2len.json", id2len), ("txt2img.json", txt2img), ("img2txts.json", img2txts),
                      ("meta.json", {"CLS": cls_id, "SEP": sep_id, "MASK": mask_id, "v_range": list(v_range)})):
        with open(os.path.join(db_dir, name), "w") as f:
            json.dump(obj, f)
    return db_dir


def write_img_db(img_dir, features, conf_th=0.2, max_bb=100, min_bb=10, num_bb=36, compress=True, backend=None):
    """Write a region-feature database (scripts/convert_imgdir.py:29-115 [layout]): features = {file name: dict with
    'features' [n, 2048], 'norm_bb' [n, 6] and 'conf' [n]} (stored fp16 like the reference's converter).  Produces
    <img_dir>/<db name>/ and, for a confidence-threshold database, nbb_th*_max*_min*.json."""
    name = _img_db_name(conf_th, max_bb, min_bb, num_bb, compress, True)
    kv = open_kv(os.path.join(img_dir, name), write=True, backend=backend)
    name2nbb = {}
    for fname, arrs in features.items():
        stored = {k: (np.asarray(v).astype(np.float16) if np.asarray(v).dtype.kind == "f" else np.asarray(v))
                  for k, v in arrs.items()}
        kv.put(fname.encode("utf-8"), pack_features(stored, compress))
        if conf_th != -1:
            name2nbb[fname] = compute_num_bb(arrs['conf'], conf_th, min_bb, max_bb)
    kv.put(b"__keys__", json.dumps(list(features.keys())).encode("utf-8"))
    kv.close()
    if conf_th != -1:
        with open(os.path.join(img_dir, f"nbb_th{conf_th}_max{max_bb}_min{min_bb}.json"), "w") as f:
            json.dump(name2nbb, f)
    return img_dir
