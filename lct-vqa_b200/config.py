"""Run-time configuration read by the modules at call time (mirrors darts_vqa/config.py:1-8 and the
globals of basic_vqa/config.py that the hot path reads)."""
import torch

# device every module is created on / moved to (darts_vqa/config.py:4, basic_vqa/config.py:56)
DEVICE = 'cuda' if torch.cuda.is_available() else 'cpu'
# seed (darts_vqa/config.py:6)
SEED = 10
ROOT_STATS_DIR = './experiment_data'

# --- basic_vqa (LCT) knobs read by the architects / models_lct (basic_vqa/config.py) ---
MAX_QST_LEN = 30
IMG_EMBED_SIZE = 512
WORD_EMBED_SIZE = 300
LSTM_NUM_LAYERS = 1
LSTM_HIDDEN_SIZE = 512
LEARNING_RATE = 0.001
ARCH_LEARNING_RATE = 6e-4     # basic_vqa/config.py:34
ARCH_WEIGHT_DECAY = 1e-3      # basic_vqa/config.py:36
GRAD_CLIP = 5
TEMPERATURE = 0.1             # basic_vqa/config.py:40
BATCH_SIZE = 64
ARCH_TYPE = 'darts'
SKIP_STAGE2 = False
SKIP_STAGE3 = True
W_LAMBDA = 1.0                # basic_vqa/config.py:74  weight of the pseudo-QA soft loss of the W model
PRETRAIN_ENC = True           # basic_vqa model_factory: VGG19 of the W model / fixed EF encoder starts from ImageNet weights
