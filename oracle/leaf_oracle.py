"""CPU ORACLE for the LEAF attack path - TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the algorithm of the reference's inner attack loop so that
the CUDA path can be checked against it. It is imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs. The product
(leaf_b200/) never imports it and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"): the reference holds NO golden vectors for this path
(SURVEY.md 8c). The pins are outputs of the reference itself, generated in the build container
by oracle/make_golden.py (which imports /root/reference unmodified through oracle/ref_shims)
and committed under tests/golden/. tests/test_oracle_golden.py checks every function below
against those fixtures. The constrain=True filter (utils_attacks.py:110-143) needs NLTK corpora
that do not exist offline: that branch is "parity unpinned" and enters only as a host-supplied
valid[B][n] mask.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import html
import os
import string

import numpy as np
import regex
import torch
import torch.nn.functional as F

SOT = 49406
EOT = 49407
CONTEXT_LENGTH = 77

# train_AT_text_only.py:93
V_DEFAULT = [-1] + [ord(c) for c in string.ascii_lowercase + " " + string.ascii_uppercase
                    + string.digits + string.punctuation]

_MERGES_BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "leaf_b200", "data",
                           "clip_bpe_merges.bin")


# ----------------------------------------------------------------------------------------------
# candidate expansion
# ----------------------------------------------------------------------------------------------
def edit_sentence(S: str, z: int, c: int) -> str:
    """One Levenshtein-1 edit; closed form of generate_sentence(S, z, u, V, k=1, alternative=-1)
    with c = V[u] (utils_attacks.py:169-213).

    The reference interleaves '_' placeholders: position z=2i is the slot before character i,
    z=2i+1 is character i. Writing '_' (c == -1, or c equal to what is already there under
    alternative == -1) clears the mask, i.e. deletes a character / leaves a slot empty.
    """
    i = z // 2
    if z % 2 == 1:                       # a character
        if c == -1 or chr(c) == S[i]:
            return S[:i] + S[i + 1:]
        return S[:i] + chr(c) + S[i + 1:]
    if c == -1 or chr(c) == "_":         # a slot: placeholder already holds '_'
        return S
    return S[:i] + chr(c) + S[i:]


def expand_positions(S: str, positions) -> list:
    """Phase-1 probe set: a space at each drawn position
    (generate_all_sentences(S, [ord(' ')], subset_z=positions, alternative=-1),
    utils_attacks.py:275-295,318)."""
    return [edit_sentence(S, int(z), ord(" ")) for z in positions]


def expand_chars(S: str, z: int, V, us) -> list:
    """Phase-2 set: characters V[u] at the chosen position
    (generate_random_sentences_at_z, utils_attacks.py:226-236)."""
    return [edit_sentence(S, int(z), V[int(u)]) for u in us]


def draw_positions(S: str, n: int):
    """utils_attacks.py:317 - consumes the global numpy RNG exactly as the reference does."""
    return np.random.choice(range(2 * len(S) + 1), size=n, replace=n > 2 * len(S) + 1)


def draw_chars(V, n: int):
    """utils_attacks.py:236."""
    return np.random.choice(range(len(V)), size=n, replace=(n > len(V)))


# ----------------------------------------------------------------------------------------------
# tokenizer
# ----------------------------------------------------------------------------------------------
def byte_symbol_ids():
    """id of each byte's symbol = its index in bytes_to_unicode() order (tokenizer.py:31-51,147)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    for b in range(256):
        if b not in bs:
            bs.append(b)
    ids = [0] * 256
    for i, b in enumerate(bs):
        ids[b] = i
    return ids


class OracleTokenizer:
    """Restatement of SimpleTokenizer (tokenizer.py:133-265) on integer symbol ids.

    Same third-party/stdlib calls as the reference for cleaning and splitting (`html.unescape`,
    `regex`); `ftfy.fix_text` is the identity here as in the golden generator (ftfy is not
    installed; exact for the ASCII parity domain). BPE runs on ids with the rank table
    (merged id = 512 + rank)."""

    PAT = regex.compile(
        r"""<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
        regex.IGNORECASE)                                              # tokenizer.py:160-163

    # transformers CLIPTokenizer (models/clip/tokenization_clip.py): same BPE, but no html.unescape and the special
    # tokens are spelled <|startoftext|> / <|endoftext|>
    PAT_HF = regex.compile(
        r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
        regex.IGNORECASE)

    def __init__(self, merges_path: str = _MERGES_BIN, context_length: int = CONTEXT_LENGTH, hf: bool = False):
        self.hf = hf
        pairs = np.fromfile(merges_path, dtype="<u4")
        assert pairs.shape[0] == 48894
        self.ranks = {(int(p) >> 16, int(p) & 0xFFFF): r for r, p in enumerate(pairs)}
        self.byte_id = byte_symbol_ids()
        self.context_length = context_length
        self.cache = {}

    # tokenizer.py:66-69,72-75,83-85
    @staticmethod
    def clean(text: str) -> str:
        text = html.unescape(html.unescape(text)).strip()
        text = " ".join(text.split()).strip()
        return text.lower()

    def bpe(self, ids):
        """tokenizer.py:172-211 on ids: repeatedly merge the lowest-rank adjacent pair, every
        occurrence left to right."""
        word = list(ids)
        while len(word) > 1:
            best = None
            for a, b in zip(word[:-1], word[1:]):
                r = self.ranks.get((a, b))
                if r is not None and (best is None or r < best[0]):
                    best = (r, a, b)
            if best is None:
                break
            r, a, b = best
            new, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    new.append(512 + r)
                    i += 2
                else:
                    new.append(word[i])
                    i += 1
            word = new
        return word

    def encode(self, text: str):
        """tokenizer.py:213-219. hf=True: CLIPTokenizer._tokenize without ftfy (BasicTokenizer: whitespace clean + lower,
        which is all it does to printable ASCII), HF's pattern and special-token spellings."""
        out = []
        cleaned = " ".join(text.split()).strip().lower() if self.hf else self.clean(text)
        sot, eot = ("<|startoftext|>", "<|endoftext|>") if self.hf else ("<start_of_text>", "<end_of_text>")
        for piece in (self.PAT_HF if self.hf else self.PAT).findall(cleaned):
            if piece == sot:                     # self.cache seeds the specials (tokenizer.py:159)
                out.append(SOT)
                continue
            if piece == eot:
                out.append(EOT)
                continue
            got = self.cache.get(piece)
            if got is None:
                sym = [self.byte_id[b] for b in piece.encode("utf-8")]
                sym[-1] += 256                   # last symbol carries '</w>' (tokenizer.py:175)
                got = self.bpe(sym)
                self.cache[piece] = got
            out.extend(got)
        return out

    def __call__(self, texts, context_length=None):
        """tokenizer.py:226-265 - returns int64 [N, ctx], zero padded, truncated with EOT forced."""
        if isinstance(texts, str):
            texts = [texts]
        ctx = context_length or self.context_length
        res = np.zeros((len(texts), ctx), dtype=np.int64)
        for i, t in enumerate(texts):
            toks = [SOT] + self.encode(t) + [EOT]
            if len(toks) > ctx:
                toks = toks[:ctx]
                toks[-1] = EOT
            res[i, :len(toks)] = toks
        return torch.from_numpy(res)

    def hf_call(self, texts, pad_id: int = EOT, max_length: int = CONTEXT_LENGTH):
        """tokenizer_wrapper.__call__ (utils_attacks.py:67-71): tokenizer(x, padding=True, truncation=True).input_ids as a
        tensor - padded to the LONGEST row of the batch with the tokenizer's pad id, truncated to max_length with the
        EOS kept."""
        assert self.hf
        rows = []
        for t in ([texts] if isinstance(texts, str) else texts):
            toks = [SOT] + self.encode(t) + [EOT]
            if len(toks) > max_length:
                toks = toks[:max_length - 1] + [EOT]
            rows.append(toks)
        L = max(len(r) for r in rows)
        return torch.tensor([r + [pad_id] * (L - len(r)) for r in rows], dtype=torch.int64)


# ----------------------------------------------------------------------------------------------
# text tower (fp32)
# ----------------------------------------------------------------------------------------------
def _act(x, quick_gelu: bool):
    if quick_gelu:
        return x * torch.sigmoid(1.702 * x)        # transformer.py:33-36
    return F.gelu(x)                               # nn.GELU (erf), model.py:192


def encode_text(sd: dict, tokens: torch.Tensor, heads: int, quick_gelu: bool = False,
                normalize: bool = False, prefix: str = "") -> torch.Tensor:
    """CLIP.encode_text (model.py:269-284) over an open_clip state dict, in plain fp32 torch ops.

    x0 = tok_emb[ids] + pos_emb; L x pre-LN blocks (transformer.py:254-265) with causal additive
    mask (-inf above the diagonal, transformer.py:758-764), head_dim = W/heads, scale 1/sqrt(d);
    ln_final; pool at argmax(ids) (transformer.py:661); @ text_projection."""
    g = lambda k: sd[prefix + k].float()
    tokens = tokens.long()
    N, T = tokens.shape
    x = g("token_embedding.weight")[tokens] + g("positional_embedding")[:T]
    W = x.shape[-1]
    d = W // heads
    mask = torch.full((T, T), float("-inf")).triu_(1)
    i = 0
    while (prefix + f"transformer.resblocks.{i}.ln_1.weight") in sd:
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (W,), g(p + "ln_1.weight"), g(p + "ln_1.bias"), 1e-5)
        qkv = h @ g(p + "attn.in_proj_weight").T + g(p + "attn.in_proj_bias")
        q, k, v = qkv.split(W, dim=-1)
        q = q.view(N, T, heads, d).transpose(1, 2)
        k = k.view(N, T, heads, d).transpose(1, 2)
        v = v.view(N, T, heads, d).transpose(1, 2)
        att = (q @ k.transpose(-1, -2)) * (d ** -0.5) + mask
        att = torch.softmax(att, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(N, T, W)
        x = x + o @ g(p + "attn.out_proj.weight").T + g(p + "attn.out_proj.bias")
        h = F.layer_norm(x, (W,), g(p + "ln_2.weight"), g(p + "ln_2.bias"), 1e-5)
        h = _act(h @ g(p + "mlp.c_fc.weight").T + g(p + "mlp.c_fc.bias"), quick_gelu)
        x = x + h @ g(p + "mlp.c_proj.weight").T + g(p + "mlp.c_proj.bias")
        i += 1
    x = F.layer_norm(x, (W,), g("ln_final.weight"), g("ln_final.bias"), 1e-5)
    pooled = x[torch.arange(N), tokens.argmax(dim=-1)]
    out = pooled @ g("text_projection")
    return F.normalize(out, dim=-1) if normalize else out


def encode_text_device(sd: dict, tokens: torch.Tensor, heads: int, quick_gelu: bool = False) -> torch.Tensor:
    """encode_text above without the .float() copies and on the parameters' device, so that torch autograd can
    differentiate it (reference gradients for the K4 tests; same lines of model.py / transformer.py)."""
    tokens = tokens.long()
    N, T = tokens.shape
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"][:T]
    W = x.shape[-1]
    d = W // heads
    mask = torch.full((T, T), float("-inf"), device=x.device).triu_(1)
    i = 0
    while f"transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (W,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = h @ sd[p + "attn.in_proj_weight"].T + sd[p + "attn.in_proj_bias"]
        q, k, v = (z.view(N, T, heads, d).transpose(1, 2) for z in qkv.split(W, dim=-1))
        att = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5) + mask, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(N, T, W)
        x = x + o @ sd[p + "attn.out_proj.weight"].T + sd[p + "attn.out_proj.bias"]
        h = F.layer_norm(x, (W,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = _act(h @ sd[p + "mlp.c_fc.weight"].T + sd[p + "mlp.c_fc.bias"], quick_gelu)
        x = x + h @ sd[p + "mlp.c_proj.weight"].T + sd[p + "mlp.c_proj.bias"]
        i += 1
    x = F.layer_norm(x, (W,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(N, device=x.device), tokens.argmax(dim=-1)] @ sd["text_projection"]


# ----------------------------------------------------------------------------------------------
# loss / argmax and the attack driver
# ----------------------------------------------------------------------------------------------
def score(text_features: torch.Tensor, anchor: torch.Tensor, objective: str = "l2") -> torch.Tensor:
    """utils_attacks.py:332-346 - [B,n,E],[B,E] -> [B,n]."""
    B = anchor.shape[0]
    if objective == "l2":
        return ((text_features - anchor.view(B, 1, -1)) ** 2).sum(dim=-1)
    if objective == "negl2":
        return -((text_features - anchor.view(B, 1, -1)) ** 2).sum(dim=-1)
    if objective == "dissim":
        return -(text_features @ anchor.view(B, -1, 1)).squeeze(-1)
    if objective == "sim":
        return (text_features @ anchor.view(B, -1, 1)).squeeze(-1)
    raise ValueError(objective)


def attack_text_leaf_oracle(encode, tokenizer, sentences, anchor_features, objective="l2", n=10, k=1,
                            V=V_DEFAULT, valid_fn=None, trace=None):
    """attack_text_leaf (utils_attacks.py:297-393) with `encode(tokens, normalize) -> [N,E]`.

    Draws from the global numpy RNG in the reference's order (B position draws, then B character
    draws, per round). `valid_fn(sentences, SS) -> bool[B][n]` stands in for valid_sentence_batched
    (utils_attacks.py:321-325,360-364); None == constrain=False. `trace`, if a dict, receives the
    per-round positions / characters / scores for the tests.

    NOTE the reference's dissim/sim branches use `anchor.transpose(-1,-2)` on a [B,E] anchor, which
    only type-checks for B == 1 (eval_textfare.py:127); `score` keeps that meaning per sample."""
    sentences = list(sentences)
    B = len(sentences)
    anchor = anchor_features
    if objective in ("dissim", "sim"):
        anchor = anchor / anchor.norm(dim=-1, keepdim=True)          # :304-308 (in place there)
    norm = objective in ("sim", "dissim")
    feats = ids_best = None
    for rnd in range(k):
        positions = [draw_positions(S, n) for S in sentences]                      # :317
        SS = [expand_positions(S, positions[i]) for i, S in enumerate(sentences)]   # :318
        if valid_fn is not None:                                                    # :321-325
            valid = valid_fn(sentences, SS)
            SS = [[SS[i][j] if valid[i][j] else sentences[i] for j in range(n)] for i in range(B)]
        flat = [s for sub in SS for s in sub]
        feats = encode(tokenizer(flat), norm).view(B, n, -1)                        # :327-330
        loss1 = score(feats, anchor, objective)
        ids_best = torch.argmax(loss1, dim=-1)                                      # :348
        best_pos = [int(positions[r][int(i)]) for r, i in enumerate(ids_best)]      # :350-353
        us = [draw_chars(V, n) for _ in sentences]                                  # :355-357 -> :236
        SS = [expand_chars(S, best_pos[i], V, us[i]) for i, S in enumerate(sentences)]
        if valid_fn is not None:                                                    # :360-364
            valid = valid_fn(sentences, SS)
            SS = [[SS[i][j] if valid[i][j] else sentences[i] for j in range(n)] for i in range(B)]
        flat = [s for sub in SS for s in sub]
        feats = encode(tokenizer(flat), norm).view(B, n, -1)                        # :366-368
        loss2 = score(feats, anchor, objective)
        ids_best = torch.argmax(loss2, dim=-1)                                      # :386
        if trace is not None:
            trace.setdefault("rounds", []).append(dict(
                sentences=list(sentences), positions=[p.tolist() for p in positions],
                best_pos=best_pos, chars=[[V[int(u)] for u in uu] for uu in us],
                loss1=loss1.clone(), loss2=loss2.clone(), ids_best=ids_best.clone()))
        sentences = [flat[r * n + int(i)] for r, i in enumerate(ids_best)]          # :387-389
    best = torch.take_along_dim(feats, ids_best.view(-1, 1, 1).repeat(1, 1, feats.shape[-1]),
                                dim=1).squeeze(1)                                   # :393
    return best, sentences


# ----------------------------------------------------------------------------------------------
# single-sentence evaluation attacks
# ----------------------------------------------------------------------------------------------
def all_sentences(S: str, V, subset_z=None) -> list:
    """generate_all_sentences(S, V, subset_z, k=1, alternative=-1) (utils_attacks.py:275-295): for every position z of
    subset_z (default: all 2*len(S)+1), every character of V."""
    if subset_z is None:
        subset_z = range(2 * len(S) + 1)
    return [edit_sentence(S, int(z), V[u]) for z in subset_z for u in range(len(V))]


def _batched_loss(encode, tokenizer, SS, anchor, objective, batch_size, encode_2=None, anchor_2=None):
    """The reference's batching loop (utils_attacks.py:420-445, :483-517, :538-573) INCLUDING its off-by-one:
    `end = min((i+1)*bs, len-1)` never evaluates the LAST candidate. Returns loss[len-1]."""
    tokens = tokenizer(SS)
    norm = objective in ("sim", "dissim")
    out = []
    for i in range(len(tokens) // batch_size + 1):
        beg, end = i * batch_size, min((i + 1) * batch_size, len(tokens) - 1)
        if beg >= end:
            continue
        f = encode(tokens[beg:end], norm).view(end - beg, -1)
        l = score(f.unsqueeze(0), anchor.view(1, -1), objective).squeeze(0)
        if encode_2 is not None:
            f2 = encode_2(tokens[beg:end], norm).view(end - beg, -1)
            l = (l + score(f2.unsqueeze(0), anchor_2.view(1, -1), objective).squeeze(0)) / 2
        out.append(l)
    return torch.cat(out, dim=0)


def attack_text_charmer_oracle(encode, tokenizer, sentence, anchor_features, objective="l2", n=10, k=1, V=V_DEFAULT,
                               valid_fn=None, batch_size=20 * 128, encode_2=None, anchor_2=None, trace=None):
    """attack_text_charmer_inference (utils_attacks.py:451-580): per round, probe every position with a space, keep the
    top-n positions (torch.topk), then try every character of V at those positions and keep the argmax.
    Deviation noted in DESIGN.md: with model_2 and objective 'l2' the reference's first phase raises (a misplaced `/2`
    on list.append's None, :496); here every objective averages the two losses, which is what :498-513 do."""
    anchor = anchor_features
    if objective in ("dissim", "sim"):
        anchor = anchor / anchor.norm(dim=-1, keepdim=True)
        if anchor_2 is not None:
            anchor_2 = anchor_2 / anchor_2.norm(dim=-1, keepdim=True)
    dist = 0
    for dist in range(k):
        SS = all_sentences(sentence, [ord(" ")])                                   # :476-477
        if valid_fn is not None:                                                   # :478-481
            valid = valid_fn([sentence], [SS])[0]
            SS = [s if v else sentence for s, v in zip(SS, valid)]
        loss = _batched_loss(encode, tokenizer, SS, anchor, objective, batch_size, encode_2, anchor_2)
        top = torch.topk(loss, min(n, loss.shape[0]), dim=0).indices               # :519
        SS = all_sentences(sentence, V, subset_z=[int(z) for z in top])            # :524
        if valid_fn is not None:                                                   # :532-537
            valid = valid_fn([sentence], [SS])[0]
            SS = [s if v else sentence for s, v in zip(SS, valid)]
        loss2 = _batched_loss(encode, tokenizer, SS, anchor, objective, batch_size, encode_2, anchor_2)
        if trace is not None:
            trace.setdefault("rounds", []).append(dict(sentence=sentence, loss1=loss.clone(), top=top.clone(), loss2=loss2.clone()))
        sentence = SS[int(torch.argmax(loss2))]                                    # :575
    return sentence, dist + 1


def attack_text_bruteforce_oracle(encode, tokenizer, sentence, anchor_features, objective="l2", V=V_DEFAULT,
                                  valid_fn=None, batch_size=20 * 128, trace=None):
    """attack_text_bruteforce (utils_attacks.py:395-449): every position x every character, k = 1; objectives 'l2' and
    'dissim' only (anything else leaves the loss undefined there and raises)."""
    if objective not in ("l2", "dissim"):
        raise ValueError(objective)
    anchor = anchor_features
    if objective == "dissim":
        anchor = anchor / anchor.norm(dim=-1, keepdim=True)
    SS = all_sentences(sentence, V)                                                # :412
    if valid_fn is not None:                                                       # :415-418
        valid = valid_fn([sentence], [SS])[0]
        SS = [s if v else sentence for s, v in zip(SS, valid)]
    loss = _batched_loss(encode, tokenizer, SS, anchor, objective, batch_size)
    if trace is not None:
        trace["loss"] = loss.clone()
    return SS[int(torch.argmax(loss))], 1
