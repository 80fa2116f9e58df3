"""Host logic of leaf_b200.attack_text_leaf on the CPU with the oracle-backed engine double (tests/oracle_engine.py):
the reference's draw order and phase chaining against the golden attack runs, and the two sharding modes under a
world_size-2 gloo group against the unsharded result."""
import json
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leaf_b200 import attack_text_leaf, synth
from leaf_b200 import dist as D
from tests.oracle_engine import OracleEngine

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _setup():
    g = json.load(open(os.path.join(GOLDEN, "attack_golden.json")))
    z = np.load(os.path.join(GOLDEN, "attack_golden.npz"))
    cfg = synth.TOWERS[g["tower"]]
    sd = synth.random_tower_state_dict(cfg, seed=g["seed"], exact_numpy=True)
    return g, z, cfg, sd


def test_attack_driver_reproduces_reference_runs():
    g, z, cfg, sd = _setup()
    eng = OracleEngine(sd, cfg.heads)
    for ci, c in enumerate(g["cases"]):
        np.random.seed(c["seed"])
        feats, adv = attack_text_leaf(eng, None, c["captions"], torch.from_numpy(z[f"anchor_{ci}"]).clone(), "cpu",
                                      objective=c["objective"], n=c["n"], k=c["k"])
        assert adv == c["adv"], ci
        assert np.abs(feats.numpy() - z[f"feats_{ci}"]).max() < 1e-4


def test_attack_driver_on_accented_captions():
    """The same on captions with code points up to U+024F: UTF-8 on the wire, positions in code points, the K1 core's wider domain."""
    g = json.load(open(os.path.join(GOLDEN, "attack_latin_golden.json")))
    z = np.load(os.path.join(GOLDEN, "attack_latin_golden.npz"))
    cfg = synth.TOWERS[g["tower"]]
    eng = OracleEngine(synth.random_tower_state_dict(cfg, seed=g["seed"], exact_numpy=True), cfg.heads)
    for ci, c in enumerate(g["cases"]):
        np.random.seed(c["seed"])
        feats, adv = attack_text_leaf(eng, None, c["captions"], torch.from_numpy(z[f"anchor_{ci}"]).clone(), "cpu", n=c["n"], k=c["k"])
        assert adv == c["adv"], ci
        assert np.abs(feats.numpy() - z[f"feats_{ci}"]).max() < 1e-4


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def test_eval_attack_drivers_reproduce_reference_runs():
    """leaf_b200.attack_text_charmer_inference / attack_text_bruteforce (host logic over the engine double) against
    runs of the reference's own functions (utils_attacks.py:395-580), including the never-scored last candidate."""
    from leaf_b200 import attack_text_bruteforce, attack_text_charmer_inference
    g = json.load(open(os.path.join(GOLDEN, "eval_attack_golden.json")))
    z = np.load(os.path.join(GOLDEN, "eval_attack_golden.npz"))
    cfg = synth.TOWERS[g["tower"]]
    eng = OracleEngine(synth.random_tower_state_dict(cfg, seed=g["seed"], exact_numpy=True), cfg.heads)
    eng2 = OracleEngine(synth.random_tower_state_dict(cfg, seed=g["seed2"], exact_numpy=True), cfg.heads)
    for ci, c in enumerate(g["charmer"]):
        a2 = torch.from_numpy(z[f"charmer_anchor2_{ci}"]).clone() if c["two"] else None
        adv, dist_ = attack_text_charmer_inference(eng, None, c["sentence"], torch.from_numpy(z[f"charmer_anchor_{ci}"]).clone(),
                                                   "cpu", objective=c["objective"], n=c["n"], k=c["k"],
                                                   batch_size=c["batch_size"], model_2=eng2 if c["two"] else None,
                                                   model_2_anchor_features=a2)
        if c["tie_dependent"]:        # the reference's own result hangs on torch.topk's unspecified order of equal scores
            assert dist_ == c["dist"] and _levenshtein(adv, c["sentence"]) <= c["k"], ci
        else:
            assert (adv, dist_) == (c["adv"], c["dist"]), ci
    for ci, c in enumerate(g["bruteforce"]):
        adv, dist_ = attack_text_bruteforce(eng, None, c["sentence"], torch.from_numpy(z[f"brute_anchor_{ci}"]).clone(), "cpu",
                                            batch_size=c["batch_size"], objective=c["objective"])
        assert (adv, dist_) == (c["adv"], c["dist"]), ci
    # the constraint mask: nothing valid -> the sentence comes back unchanged (utils_attacks.py:478-481, :532-537)
    nothing_valid = lambda sentences, SS: [[False] * len(SS[0]) for _ in sentences]
    c = g["charmer"][0]
    adv, _ = attack_text_charmer_inference(eng, None, c["sentence"], torch.from_numpy(z["charmer_anchor_0"]).clone(), "cpu",
                                           n=c["n"], k=1, constrain=nothing_valid)
    assert adv == c["sentence"]


def test_constraint_mask_replaces_candidates_by_the_current_sentence():
    g, z, cfg, sd = _setup()
    eng = OracleEngine(sd, cfg.heads)
    c = g["cases"][0]
    np.random.seed(c["seed"])
    nothing_valid = lambda sentences, SS: [[False] * len(SS[0]) for _ in sentences]
    _, adv = attack_text_leaf(eng, None, c["captions"], torch.from_numpy(z["anchor_0"]).clone(), "cpu", n=c["n"], k=1,
                              constrain=nothing_valid)
    assert adv == c["captions"]                       # utils_attacks.py:325/:364: invalid -> the sentence itself


def test_device_constraint_reproduces_the_oracle_attack_with_the_filter():
    """constrain=True (masks from the filter core, no candidate strings on the host) against the oracle attack loop driven
    by the oracle's restatement of valid_sentence_batched, k = 2 so that round 2 compares against round 1's winners."""
    from oracle import leaf_oracle as O
    from oracle import nltk_restate as N
    from tests.test_constrain_cpu import ABBREV, _word_list
    g, z, cfg, sd = _setup()
    words = _word_list()
    eng = OracleEngine(sd, cfg.heads)
    eng.load_words(words, ABBREV)
    W, A = frozenset(words), frozenset(ABBREV)
    caps = synth.make_captions(5, seed=9)
    anchor = torch.from_numpy(z["anchor_0"])[:5].clone()
    otok = O.OracleTokenizer()
    enc = lambda t, normalize: O.encode_text(sd, t, cfg.heads, normalize=normalize)
    np.random.seed(5)
    want_f, want = O.attack_text_leaf_oracle(enc, otok, caps, anchor, n=30, k=2,
                                             valid_fn=lambda s, SS: N.valid_sentence_batched(s, SS, W, A))
    np.random.seed(5)
    feats, adv = attack_text_leaf(eng, None, caps, anchor.clone(), "cpu", n=30, k=2, constrain=True)
    assert adv == want
    assert np.abs(feats.numpy() - want_f.numpy()).max() < 1e-4
    assert adv != caps


def test_cross_shard_argmax_single_process():
    v, i = torch.tensor([1.0, 2.0]), torch.tensor([3, 4])
    assert D.cross_shard_argmax(v, i) == (v, i)
    assert D.shard_range(10, 0, 3) == (0, 3) and D.shard_range(10, 2, 3) == (6, 10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g, z, cfg, sd = _setup()
    res = {}
    # first-index tie-break across shards: ties must resolve to the smallest GLOBAL index
    val = torch.tensor([5.0, 1.0, 7.0]) if rank == 0 else torch.tensor([5.0, 2.0, 7.0])
    idx = torch.tensor([4, 0, 9]) if rank == 0 else torch.tensor([2, 6, 11])
    gv, gi = D.cross_shard_argmax(val, idx)
    res["argmax"] = (gv.tolist(), gi.tolist())
    # a NaN score never wins over a number, an empty shard (NO_CANDIDATE) never wins at all; the packed pair survives the
    # round trip bit for bit (negative scores, -0.0, indices up to 2^31 - 2)
    val = torch.tensor([float("nan"), -3.5, float("-inf")]) if rank == 0 else torch.tensor([1.0, float("-inf"), float("nan")])
    idx = torch.tensor([0, 7, 3]) if rank == 0 else torch.tensor([5, D.NO_CANDIDATE, 2])
    gv, gi = D.cross_shard_argmax(val, idx)
    res["argmax_nan"] = (gv.tolist(), gi.tolist())
    pv, pi = torch.tensor([-0.0, -1.25e-30, 3.4e38, float("inf")]), torch.tensor([0, 2 ** 31 - 2, 49, D.NO_CANDIDATE])
    uv, ui = D.unpack_score_index(D.pack_score_index(pv, pi))
    res["pack"] = bool(torch.equal(uv.view(torch.int32), pv.view(torch.int32)) and torch.equal(ui, pi))
    # broadcast_rows selects the owner's rows: NaN / Inf in rows a rank does not own must not leak
    rows = torch.tensor([[1.0, 2.0], [float("nan"), float("inf")]]) if rank == 0 else torch.tensor([[float("inf"), float("nan")], [3.0, 4.0]])
    res["bcast"] = D.broadcast_rows(rows, torch.tensor([0, 1])).tolist()
    # fewer candidates than ranks (n = 1 < world size): rank 1 scores nothing and must still take part in every collective
    eng = OracleEngine(sd, cfg.heads)
    c = g["cases"][0]
    np.random.seed(77)
    _, adv1 = attack_text_leaf(eng, None, c["captions"][:3], torch.from_numpy(z["anchor_0"])[:3].clone(), "cpu", n=1, k=2,
                               shard="candidates")
    np.random.seed(77)
    _, adv0 = attack_text_leaf(OracleEngine(sd, cfg.heads), None, c["captions"][:3], torch.from_numpy(z["anchor_0"])[:3].clone(),
                               "cpu", n=1, k=2)
    res["n_lt_world"] = (adv1 == adv0, eng.encoded_rows)
    # fewer samples than ranks (B = 1): rank 1 attacks nothing in the sample-sharded mode
    np.random.seed(78)
    _, adv1 = attack_text_leaf(OracleEngine(sd, cfg.heads), None, c["captions"][:1], torch.from_numpy(z["anchor_0"])[:1].clone(),
                               "cpu", n=6, k=1, shard="samples")
    np.random.seed(78)
    _, adv0 = attack_text_leaf(OracleEngine(sd, cfg.heads), None, c["captions"][:1], torch.from_numpy(z["anchor_0"])[:1].clone(),
                               "cpu", n=6, k=1)
    res["b_lt_world"] = adv1 == adv0
    for mode in ("samples", "candidates"):
        for ci in (1, 3):                             # k=1 n=50 and k=2 n=20
            c = g["cases"][ci]
            eng = OracleEngine(sd, cfg.heads)
            np.random.seed(c["seed"])
            feats, adv = attack_text_leaf(eng, None, c["captions"], torch.from_numpy(z[f"anchor_{ci}"]).clone(), "cpu",
                                          objective=c["objective"], n=c["n"], k=c["k"], shard=mode)
            res[(mode, ci)] = (adv, float(np.abs(feats.numpy() - z[f"feats_{ci}"]).max()), eng.encoded_rows)
    out[rank] = res
    dist.destroy_process_group()


def test_sharded_modes_equal_the_unsharded_run_gloo_world2():
    g, z, cfg, sd = _setup()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in (0, 1):
        res = out[rank]
        assert res["argmax"] == ([5.0, 2.0, 7.0], [2, 6, 9])
        assert res["argmax_nan"][1] == [5, 7, 3] and res["argmax_nan"][0][:2] == [1.0, -3.5]
        assert res["pack"]
        assert res["bcast"] == [[1.0, 2.0], [3.0, 4.0]]
        assert res["n_lt_world"][0] and res["b_lt_world"]
        for mode in ("samples", "candidates"):
            for ci in (1, 3):
                adv, err, rows = res[(mode, ci)]
                assert adv == g["cases"][ci]["adv"], (rank, mode, ci)
                assert err < 1e-4
    # each rank really did only its share of the work
    c = g["cases"][1]
    full = 2 * c["k"] * (c["B"] * c["n"] + c["B"])
    assert out[0][("samples", 1)][2] + out[1][("samples", 1)][2] == full
    assert out[0][("candidates", 1)][2] < 0.6 * full
    assert out[0]["n_lt_world"][1] == 0 and out[1]["n_lt_world"][1] > 0      # n = 1: rank 0 held the empty shard [0, 0)


def test_fast_draw_is_np_random_choice():
    """attack_text_leaf draws with permutation / randint instead of np.random.choice(range(N), n, replace=n > N): same numbers,
    same dtype, same state of the global stream afterwards (utils_attacks.py:317, :236) - for every population size a
    caption or the alphabet can produce."""
    from leaf_b200 import attack
    assert attack._DRAW is attack._fast_draw
    for N in list(range(1, 130)) + [161, 401, 2001]:
        for n in (1, 2, 10, 50, 200):
            np.random.seed(N * 7 + n)
            a, ra = np.random.choice(range(N), size=n, replace=n > N), np.random.random()
            np.random.seed(N * 7 + n)
            b, rb = attack._fast_draw(N, n), np.random.random()
            assert a.dtype == b.dtype and np.array_equal(a, b) and ra == rb, (N, n)
