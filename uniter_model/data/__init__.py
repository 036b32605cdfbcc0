"""uniter_model/data/__init__.py: the names eval_itm.py:15-16 / train_itm.py:18-19 import."""
from lightningdot_b200.data import (DetectFeatLmdb, DetectFeatTxtTokDataset, ImageLmdbGroup, TxtLmdb, TxtTokLmdb,  # noqa: F401
                                    get_gather_index, get_ids_and_lens, pad_tensors)
from lightningdot_b200.loader import PrefetchLoader, move_to_cuda, record_cuda_stream  # noqa: F401
