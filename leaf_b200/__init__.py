"""leaf_b200 - B200-native engine for LEAF's inner attack loop (candidate expansion + CLIP BPE, text tower,
TextFARE score + argmax). The compute lives in lib/libleaf_b200.so (CUDA, sm_100a); importing the package does not
need a GPU, using it does - there is no CPU path."""
from .attack import V_DEFAULT, attack_text, attack_text_leaf, generate_sentence  # noqa: F401
from .eval_attacks import attack_text_bruteforce, attack_text_charmer_inference  # noqa: F401
from .engine import LeafEngine  # noqa: F401
from .tower import LeafTextTower  # noqa: F401
from ._native import LeafError  # noqa: F401
