"""K4 on the GPU: gradients of the FARE objective (utils_AT.py:317-337) through the engine's backward against torch
autograd over the fp32 oracle tower, for the open_clip and the HF parameter layouts.
Tolerances (the engine multiplies bf16 operands; the reference trains under fp16 autocast):
  * backward alone (same cotangent dL/df fed to both): relative L2 error per tensor <= 3e-2, cosine >= 0.999;
  * end to end (loss.backward()): the TextFARE cotangent 2(f - a)/B is a DIFFERENCE of embeddings, so the forward's
    bf16 error (|df|/|f| ~ 7e-3) is amplified by |f|/|f - a| before it enters the backward: cosine >= 0.995."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_grads(sd, tok, anchor, heads, quick, cotangent=None):
    """fp32 autograd reference. cotangent = None: gradients of the FARE loss; else the vector-Jacobian product."""
    from oracle import leaf_oracle as O
    p = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in sd.items()}
    f = O.encode_text_device(p, tok.cuda(), heads, quick_gelu=quick)
    loss = torch.nn.functional.mse_loss(anchor, f, reduction="none").sum(-1).mean()      # utils_AT.py:321
    if cotangent is None:
        loss.backward()
    else:
        f.backward(cotangent)
    return loss.item(), f.detach(), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("name,quick", [("small", False), ("tiny", True)])
def test_backward_matches_autograd(name, quick):
    from leaf_b200 import synth
    from leaf_b200.tower import LeafTextTower
    cfg = synth.TOWERS[name]
    sd = synth.random_tower_state_dict(cfg, seed=11, exact_numpy=True)
    tower = LeafTextTower(sd, heads=cfg.heads, quick_gelu=quick).trainable()
    caps = synth.make_captions(6, seed=5) + synth.make_captions(1, seed=5, kind="dense-77") + ["a", ""]
    tok = tower.tokenizer(caps)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        anchor = tower.encode_text(tok) + 0.3 * torch.randn((len(caps), cfg.embed_dim), generator=g).cuda()
    f = tower.encode_text(tok)
    assert f.requires_grad
    loss = torch.nn.functional.mse_loss(anchor, f, reduction="none").sum(-1).mean()
    loss.backward()
    ref_loss, ref_f, ref = _ref_grads(sd, tok, anchor, cfg.heads, quick)
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss)
    cot = (2.0 * (f.detach() - anchor) / len(caps)).contiguous()            # the cotangent the engine's backward was given
    _, _, vjp = _ref_grads(sd, tok, anchor, cfg.heads, quick, cotangent=cot)

    def err(got, want):
        rel = ((got - want).norm() / want.norm().clamp_min(1e-20)).item()
        return rel, torch.nn.functional.cosine_similarity(got.flatten().double(), want.flatten().double(), dim=0).item()

    for k, safe in tower._names.items():
        got = getattr(tower, safe).grad
        assert got is not None and got.shape == ref[k].shape, k
        rel, cos = err(got, vjp[k])
        assert rel <= 3e-2 and cos >= 0.999, ("backward", k, rel, cos)
        rel, cos = err(got, ref[k])
        assert cos >= 0.995, ("end to end", k, rel, cos)
    # a second backward accumulates (+=) like torch
    f2 = tower.encode_text(tok)
    torch.nn.functional.mse_loss(anchor, f2, reduction="none").sum(-1).mean().backward()
    k0 = "transformer.resblocks.0.mlp.c_fc.weight"
    assert err(getattr(tower, tower._names[k0]).grad, 2 * vjp[k0])[0] <= 3e-2


def test_backward_hf_layout():
    from leaf_b200 import synth
    from leaf_b200.engine import LeafEngine
    cfg = synth.TOWERS["small"]
    sd = {k: v.cuda() for k, v in synth.random_tower_state_dict(cfg, seed=4, exact_numpy=True).items()}
    W = cfg.width
    hf = {"text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
          "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
          "text_model.final_layer_norm.weight": sd["ln_final.weight"], "text_model.final_layer_norm.bias": sd["ln_final.bias"],
          "text_projection.weight": sd["text_projection"].T.contiguous()}
    for i in range(cfg.layers):
        p, q = f"transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        for j, nm in enumerate("qkv"):
            hf[q + f"self_attn.{nm}_proj.weight"] = sd[p + "attn.in_proj_weight"][j * W:(j + 1) * W].contiguous()
            hf[q + f"self_attn.{nm}_proj.bias"] = sd[p + "attn.in_proj_bias"][j * W:(j + 1) * W].contiguous()
        for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                     ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            hf[q + b + ".weight"], hf[q + b + ".bias"] = sd[p + a + ".weight"], sd[p + a + ".bias"]
    e1, e2 = LeafEngine(sd, heads=cfg.heads), LeafEngine(hf, heads=cfg.heads)
    tok = e1.tokenize(synth.make_captions(5, seed=1))
    f1, f2 = e1.forward_train(tok), e2.forward_train(tok)
    assert torch.equal(f1, f2)
    # train-mode forward == inference forward (no dropout in the text tower); the two paths round fc1's output at
    # different points (fused GELU epilogue vs saved pre-activation), hence a bf16-level tolerance
    fi = e1.encode_tokens(tok)
    assert (f1 - fi).norm() / fi.norm() < 1e-2
    d = torch.randn_like(f1)
    g1 = {k: torch.zeros_like(v) for k, v in sd.items()}
    g2 = {k: torch.zeros_like(v) for k, v in hf.items()}
    e1.backward(d, g1)
    e2.backward(d, g2)
    assert torch.allclose(g2["text_projection.weight"], g1["text_projection"].T, rtol=1e-3, atol=1e-6)
    for i in range(cfg.layers):
        p, q = f"transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        for j, nm in enumerate("qkv"):
            assert torch.allclose(g2[q + f"self_attn.{nm}_proj.weight"], g1[p + "attn.in_proj_weight"][j * W:(j + 1) * W], rtol=1e-3, atol=1e-6)
            assert torch.allclose(g2[q + f"self_attn.{nm}_proj.bias"], g1[p + "attn.in_proj_bias"][j * W:(j + 1) * W], rtol=1e-3, atol=1e-6)
        assert torch.allclose(g2[q + "mlp.fc1.weight"], g1[p + "mlp.c_fc.weight"], rtol=1e-3, atol=1e-6)
