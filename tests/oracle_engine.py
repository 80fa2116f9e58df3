"""TEST DOUBLE of leaf_b200.engine.LeafEngine built from the oracle and the CPU-compiled K1 core, so that the host
logic of attack_text_leaf (draw order, phase chaining, sharding + collectives) can run in the `-m "not gpu"` suite,
including world_size-2 gloo runs. Never imported by the product."""
import numpy as np
import torch

from oracle import leaf_oracle as O
from tests import k1_harness as H


class OracleEngine:
    def __init__(self, sd, heads, quick_gelu=False):
        self.sd, self.heads, self.quick = sd, heads, quick_gelu
        self.device = torch.device("cpu")
        self.embed_dim = int(sd["text_projection"].shape[1])
        self.encoded_rows = 0
        self.has_words = False

    def load_words(self, words, abbrev=()):
        H.load_words(words, abbrev)
        self.has_words = True

    def constrain_mask(self, caps, off, B, n, pos, chr_, sel=None, want_counts=False):
        a = lambda t: None if t is None else t.numpy()
        counts, _ = H.constrain_counts(caps, n, a(pos), a(chr_), a(sel))
        valid = (counts[:B * n].reshape(B, n) < counts[B * n:].reshape(B, 1)).astype(np.uint8)
        return torch.from_numpy(valid)

    def reserve(self, n):
        pass

    def check_status(self):
        pass

    def upload_captions(self, sentences):
        return list(sentences), None

    def expand_tokenize(self, caps, off, B, n, pos=None, chr_=None, sel=None, valid=None):
        a = lambda t: None if t is None else t.numpy()
        tok, ln, _ = H.expand_tokenize(caps, n, a(pos), a(chr_), a(sel), a(valid))
        if n > 0:
            btok, bln, _ = H.expand_tokenize(caps, 0)
            tok, ln = np.concatenate([tok, btok]), np.concatenate([ln, bln])
            base = np.concatenate([B * n + np.arange(B * n) // n, -np.ones(B, dtype=np.int64)]).astype(np.int32)
        else:
            base = -np.ones(B, dtype=np.int32)
        return torch.from_numpy(tok), torch.from_numpy(ln), torch.from_numpy(base)

    def encode_tokens(self, tok, ln=None, normalize=False, base=None, dedup=(0, 0), trim=False):
        self.encoded_rows += tok.shape[0]
        with torch.no_grad():
            return O.encode_text(self.sd, tok.long(), self.heads, quick_gelu=self.quick, normalize=normalize)

    def score(self, feats, anchor, B, n, objective="l2", want_loss=False):
        f = feats[:B * n].view(B, n, -1)
        loss = O.score(f, anchor, objective)
        best = torch.argmax(loss, dim=-1)
        bf = f[torch.arange(B), best]
        return best.to(torch.int32), bf, (loss if want_loss else None)

    def topk(self, score_a, k, m=None, score_b=None):
        a = score_a.reshape(-1)
        m = a.numel() if m is None else m
        v = a[:m] if score_b is None else (a[:m] + score_b.reshape(-1)[:m]) / 2
        order = sorted(range(m), key=lambda i: (-float(v[i]), i))[:k]       # value descending, ties by ascending index
        idx = torch.tensor(order, dtype=torch.int32)
        return idx, v[idx.long()]
