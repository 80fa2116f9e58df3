import sys, torch
sys.path.insert(0, '.')
from leaf_b200 import synth
from leaf_b200.tower import LeafTextTower
from tests.test_gpu_backward import _ref_grads
for name, quick in (("small", False), ("tiny", True)):
    cfg = synth.TOWERS[name]
    sd = synth.random_tower_state_dict(cfg, seed=11, exact_numpy=True)
    tower = LeafTextTower(sd, heads=cfg.heads, quick_gelu=quick).trainable()
    caps = synth.make_captions(6, seed=5) + synth.make_captions(1, seed=5, kind="dense-77") + ["a", ""]
    tok = tower.tokenizer(caps)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        anchor = tower.encode_text(tok) + 0.3 * torch.randn((len(caps), cfg.embed_dim), generator=g).cuda()
    f = tower.encode_text(tok)
    loss = torch.nn.functional.mse_loss(anchor, f, reduction="none").sum(-1).mean()
    loss.backward()
    ref_loss, ref_f, ref = _ref_grads(sd, tok, anchor, cfg.heads, quick)
    print(name, "loss", loss.item(), ref_loss, "feat rel", ((f.detach()-ref_f).norm()/ref_f.norm()).item())
    for k, safe in tower._names.items():
        got, want = getattr(tower, safe).grad, ref[k]
        rel = ((got - want).norm() / want.norm().clamp_min(1e-20)).item()
        cos = torch.nn.functional.cosine_similarity(got.flatten().double(), want.flatten().double(), dim=0).item()
        print(f"  {k:55s} rel={rel:.4f} cos={cos:.5f} |want|={want.norm().item():.3e}")
