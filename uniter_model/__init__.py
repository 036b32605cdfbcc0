"""`uniter_model` - namesake of the reference's vendored UNITER package, reduced to what the bi-encoder scripts import:
uniter_model.data (ImageLmdbGroup, TxtTokLmdb, DetectFeatLmdb, PrefetchLoader ...) and the KD teacher's class name
(uniter_model.model.itm).  The model code itself (uniter_model/model/{model,layer}.py) is replaced by the sm_100a tower
kernels behind lightningdot_b200.bi_encoder."""

# (the scripts import `transformers.tokenization_bert` right after this package - train_itm.py:15-21: register the alias)
from lightningdot_b200 import compat as _compat  # noqa: E402

_compat.install()
