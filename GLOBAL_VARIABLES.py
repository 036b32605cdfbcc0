"""GLOBAL_VARIABLES.py of the reference (imported by train_itm.py:17 and dvl/data/itm.py:10)."""
import os

PROJECT_FOLDER = os.path.dirname(__file__)

N_EXAMPLES_TEACHER = 10
IMG_DIM = 2048
IMG_LABEL_DIM = 1601
BUCKET_SIZE = 8192
