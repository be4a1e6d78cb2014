"""The oracle (oracle/head_oracle.py, oracle/cfg_sample.c) against the golden vectors that
tests/golden/make_golden.py produced by executing the reference's own source files."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as O


def _load_head(d, prefix=""):
    H, E, V = int(d["H"]), int(d["E"]), int(d["V"])
    head = O.VisionHead(H, E, V)
    with torch.no_grad():
        head.output_mlp_projector.weight.copy_(torch.from_numpy(d["W1"]))
        head.output_mlp_projector.bias.copy_(torch.from_numpy(d["b1"]))
        head.vision_head.weight.copy_(torch.from_numpy(d["W2"]))
        head.vision_head.bias.copy_(torch.from_numpy(d["b2"]))
    return head


@pytest.mark.parametrize("tag", ["sigmoid", "sigmoid_smooth_sft", "hinge"])
def test_simpo_oracle_matches_reference_golden(golden_dir, tag):
    d = np.load(golden_dir / "simpo_ref.npz")
    head = _load_head(d)
    hp = dict(
        beta=float(d[f"{tag}/hp/beta"]), gamma_beta_ratio=float(d[f"{tag}/hp/gamma_beta_ratio"]),
        label_smoothing=float(d[f"{tag}/hp/label_smoothing"]), sft_weight=float(d[f"{tag}/hp/sft_weight"]),
        loss_type=str(d[f"{tag}/hp/loss_type"]),
    )
    hc, hr = torch.from_numpy(d["hidden_chosen"]), torch.from_numpy(d["hidden_rejected"])
    lc, lr = torch.from_numpy(d["labels_chosen"]), torch.from_numpy(d["labels_rejected"])
    out = O.simpo_step(head, hc, hr, lc, lr, backward=True, **hp)
    B = int(d["B"])
    tol = dict(rtol=1e-5, atol=1e-6)   # fp32 contract of BASELINE.json
    np.testing.assert_allclose(float(out["loss"]), float(d[f"{tag}/loss"]), **tol)
    np.testing.assert_allclose(out["chosen_logps"].detach().numpy(), d[f"{tag}/chosen_logps"], **tol)
    np.testing.assert_allclose(out["rejected_logps"].detach().numpy(), d[f"{tag}/rejected_logps"], **tol)
    np.testing.assert_allclose(out["losses"].detach().numpy(), d[f"{tag}/losses"], **tol)
    np.testing.assert_allclose(out["chosen_rewards"].numpy(), d[f"{tag}/chosen_rewards"], **tol)
    np.testing.assert_allclose(out["per_token_logps"].detach().numpy(), d[f"{tag}/per_token_logps"], **tol)
    np.testing.assert_allclose(out["dx"][:B].numpy(), d[f"{tag}/dx_chosen"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(out["dx"][B:].numpy(), d[f"{tag}/dx_rejected"], rtol=1e-4, atol=1e-7)
    for k in ("dW1", "db1", "dW2", "db2"):
        np.testing.assert_allclose(out[k].numpy(), d[f"{tag}/{k}"], rtol=1e-4, atol=1e-7)
    # logged metrics (train.py:432-443)
    np.testing.assert_allclose(float(out["reward_accuracy"]), float(d[f"{tag}/logged/train/rewards/accuracies"]), **tol)
    np.testing.assert_allclose(float(out["reward_margin"]), float(d[f"{tag}/logged/train/rewards/margins"]), **tol)
    np.testing.assert_allclose(float(out["logits_chosen_mean"]), float(d[f"{tag}/logged/train/logits/chosen"]),
                               rtol=1e-4, atol=1e-6)
    if hp["sft_weight"] > 0:
        np.testing.assert_allclose(float(out["sft_loss"]), float(d[f"{tag}/logged/train/sft_loss"]), **tol)


def test_analytic_gradient_coefficients_match_autograd():
    """dlogits = (g/n) (onehot - softmax) reproduces autograd of the restated loss."""
    torch.manual_seed(0)
    B, T, L, H, E, V = 2, 6, 2, 16, 16, 64
    head = O.make_head(H, E, V, seed=3, w2_gain=3.0)
    hc, hr, lc, lr = O.synthetic_simpo_batch(B, T, L, H, V, seed=5)
    hp = dict(beta=2.0, gamma_beta_ratio=0.3, label_smoothing=0.05, loss_type="sigmoid")
    out = O.simpo_step(head, hc, hr, lc, lr, backward=True, **hp)
    cc, cr = O.analytic_row_coefficients(out["chosen_logps"].detach(), out["rejected_logps"].detach(), T, **hp)
    coef = torch.cat([cc, cr])                                   # [2B]
    hidden = torch.cat([hc, hr])
    labels = torch.cat([lc, lr])
    logits = head(hidden).detach()[:, :-1]
    p = logits.softmax(-1)
    lab = labels[:, 1:]
    mask = (lab != -100)
    onehot = torch.nn.functional.one_hot(lab.clamp(min=0), V).float()
    dlogits = coef[:, None, None] * (onehot - p) * mask[..., None]
    db2 = dlogits.sum((0, 1))
    np.testing.assert_allclose(db2.numpy(), out["db2"].numpy(), rtol=1e-4, atol=1e-6)


def test_cfg_oracle_matches_reference_golden(golden_dir):
    d = np.load(golden_dir / "cfg_ref.npz")
    P, V, steps = int(d["P"]), int(d["V"]), int(d["STEPS"])
    H, E = int(d["H"]), int(d["E"])
    head = O.VisionHead(H, E, V).to(torch.bfloat16)
    with torch.no_grad():
        head.output_mlp_projector.weight.copy_(O.bits_to_bf16(d["W1_bf16"]))
        head.output_mlp_projector.bias.copy_(O.bits_to_bf16(d["b1_bf16"]))
        head.vision_head.weight.copy_(O.bits_to_bf16(d["W2_bf16"]))
        head.vision_head.bias.copy_(O.bits_to_bf16(d["b2_bf16"]))
    hidden = O.bits_to_bf16(d["hidden_bf16"])
    ref_logits = O.bits_to_bf16(d["logits_bf16"])
    w, T = float(d["cfg_weight"]), float(d["temperature"])
    for s in range(steps):
        # restated head == reference head, bit for bit on the same CPU
        logits = head(hidden[s])
        assert torch.equal(logits, ref_logits[s])
        # restated merge + softmax == the probs the reference handed to torch.multinomial
        probs = O.cfg_probs(logits, w, T)
        assert torch.equal(probs, torch.from_numpy(d["probs"][s]))
        # deterministic sampler: merged logits bit-exact vs torch bf16 op-by-op, weights/Z ~ probs
        ids, merged, weights, Z = O.cfg_sample_det(logits, w, T, torch.full((P,), 0.5), merge_mode=0)
        assert torch.equal(merged, O.cfg_merged(logits, w, T).float())
        np.testing.assert_allclose((weights / Z[:, None]).numpy(), d["probs"][s], rtol=2e-6, atol=1e-12)
        gids, *_ = O.cfg_sample_det(logits, w, T, None, greedy=True)
        assert torch.equal(gids, torch.from_numpy(d["greedy"][s]))


def test_exp_det_accuracy():
    xs = np.concatenate([-np.logspace(-6, 1.9, 400), [0.0, -87.0, -100.0, -1e4]]).astype(np.float32)
    got = np.array([O.exp_det(float(x)) for x in xs])
    ref = np.exp(xs.astype(np.float64))
    ok = ref > 1e-37
    assert np.max(np.abs(got[ok] - ref[ok]) / ref[ok]) < 3e-7
    assert np.all(got[~ok] == 0.0) or np.all(got[~ok] < 1e-36)


def test_inverse_cdf_is_distributionally_multinomial():
    """the inverse-CDF sampler draws from the same categorical as torch.multinomial (chi-square, coarse bins)"""
    g = torch.Generator().manual_seed(11)
    V, P = 16384, 1
    logits = (torch.randn(2 * P, V, generator=g) * 2.0).to(torch.bfloat16)
    probs = O.cfg_probs(logits, 5.0, 1.0)[0].double()
    n = 4000
    u = torch.rand(n, generator=g)
    ids = []
    rep = logits.repeat(32, 1)
    for i in range(0, n, 32):
        out, *_ = O.cfg_sample_det(rep, 5.0, 1.0, u[i:i + 32], merge_mode=0)
        ids.append(out)
    ids = torch.cat(ids)
    # bin codes by descending probability into 8 roughly equiprobable bins
    order = torch.argsort(probs, descending=True)
    cum = torch.cumsum(probs[order], 0)
    bin_of = torch.empty(V, dtype=torch.long)
    bin_of[order] = torch.clamp((cum * 8).long(), max=7)
    exp = torch.zeros(8, dtype=torch.double).index_add_(0, bin_of, probs) * n
    obs = torch.bincount(bin_of[ids], minlength=8).double()
    chi2 = float(((obs - exp) ** 2 / exp.clamp(min=1e-9)).sum())
    assert chi2 < 30.0, chi2   # 7 dof: P(chi2 > 30) ~ 1e-4


def test_gen_img_embeds_oracle_matches_reference_golden(golden_dir):
    """next row N1: the restated gen_aligner(gen_embed(ids)) == the reference MlpProjector + nn.Embedding, bit for bit"""
    d = np.load(golden_dir / "aligner_ref.npz")
    D, CB = int(d["D"]), int(d["CB"])
    emb = torch.nn.Embedding(CB, 8).to(torch.bfloat16)
    al = O.GenAligner(8, D).to(torch.bfloat16)
    with torch.no_grad():
        emb.weight.copy_(O.bits_to_bf16(d["gen_embed_bf16"]))
        al.layers[0].weight.copy_(O.bits_to_bf16(d["wa_bf16"]))
        al.layers[0].bias.copy_(O.bits_to_bf16(d["ba_bf16"]))
        al.layers[2].weight.copy_(O.bits_to_bf16(d["wb_bf16"]))
        al.layers[2].bias.copy_(O.bits_to_bf16(d["bb_bf16"]))
        out = O.prepare_gen_img_embeds(emb, al, torch.from_numpy(d["ids"]))
    assert torch.equal(out, O.bits_to_bf16(d["out_bf16"]))


def test_clip_adamw_oracle_matches_torch_optimizer_golden(golden_dir):
    """next row N3: the op-by-op restatement equals torch.optim.AdamW + clip_grad_norm_ (the reference's optimizer,
    ospo/wrapper/train.py:108-115, ospo/utils/train.py:30,50) on the committed three-step trajectories"""
    d = np.load(golden_dir / "adamw_ref.npz")
    H, E, V = int(d["H"]), int(d["E"]), int(d["V"])
    sizes = [V * E, E * H, V, E]
    for tag in ("a", "b"):
        lr, b1, b2, eps, wd, max_norm = [float(x) for x in d[f"{tag}_hyper"]]
        p = list(torch.from_numpy(d[f"{tag}_p0"].copy()).split(sizes))
        m = [torch.zeros_like(x) for x in p]
        v = [torch.zeros_like(x) for x in p]
        for step in range(3):
            g = list(torch.from_numpy(d[f"{tag}_g{step}"].copy()).split(sizes))
            tn = O.clip_adamw_step(p, g, m, v, step + 1, lr, (b1, b2), eps, wd, max_norm)
            if max_norm > 0:
                np.testing.assert_allclose(tn, float(d[f"{tag}_norm{step}"]), rtol=1e-6)
            np.testing.assert_allclose(torch.cat(p).numpy(), d[f"{tag}_p{step + 1}"], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(torch.cat(m).numpy(), d[f"{tag}_m{step + 1}"], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(torch.cat(v).numpy(), d[f"{tag}_v{step + 1}"], rtol=1e-6, atol=1e-12)


def test_magic_number_rounding_equals_rint_for_every_bf16_logit():
    """The CUDA samplers take n = rint(t log2 e) as (y + 1.5 * 2^23) - 1.5 * 2^23 and read n as an integer from the
    mantissa bits of y + 1.5 * 2^23 (cfg_math.cuh, exp_weight2p).  Checked here in IEEE fp32 for every bf16 value of t
    (the merged logits are bf16-valued in the default merge mode), against the oracle's rintf of the clamped product."""
    import numpy as np

    bits = np.arange(1 << 16, dtype=np.uint32) << 16
    t = bits.view(np.float32)
    t = t[np.isfinite(t)]
    with np.errstate(over="ignore"):  # the largest bf16 magnitudes overflow to +-inf and are clamped like on the GPU
        y = (t * np.float32(1.4426950408889634)).astype(np.float32)
    y = np.minimum(np.maximum(y, np.float32(-1.0e4)), np.float32(1.0e4))
    magic = np.float32(12582912.0)
    ym = (y + magic).astype(np.float32)
    n_magic = (ym - magic).astype(np.float32)
    n_rint = np.rint(y).astype(np.float32)
    assert np.array_equal(n_magic, n_rint)  # (-0.0 == +0.0 compares equal: the sign of a zero n never reaches a result)
    n_int = ym.view(np.int32) - np.int32(0x4B400000)
    assert np.array_equal(n_int, n_rint.astype(np.int32))
    # the factor 2^(n - kt) built from the integer difference equals the float construction for every e in [-120, 0]
    e = np.arange(-120, 1, dtype=np.int32)
    f_int = ((e << 23) + np.int32(0x3F800000)).view(np.float32)
    assert np.array_equal(f_int, np.ldexp(np.float32(1.0), e).astype(np.float32))


def test_chunked_fp32_restatement_matches_oracle_on_cpu():
    """tests/_gpu_ref.py (the fp32 restatement the full-size GPU tests compare against) is device-agnostic: pin it to
    the oracle here, where no GPU is needed"""
    from tests._gpu_ref import simpo_step_chunked_fp32

    H, E, V, B, T, L = 96, 64, 512, 3, 16, 2
    for hp in (dict(beta=10.0, gamma_beta_ratio=0.5, label_smoothing=0.0, loss_type="sigmoid"),
               dict(beta=2.0, gamma_beta_ratio=0.3, label_smoothing=0.0, loss_type="hinge")):
        head = O.make_head(H, E, V, seed=31, w2_gain=2.0)
        hc, hr, lc, lr = O.synthetic_simpo_batch(B, T, L, H, V, seed=32)
        ref = O.simpo_step(head, hc, hr, lc, lr, backward=True, **hp)
        w = [t.detach() for t in (head.output_mlp_projector.weight, head.output_mlp_projector.bias,
                                  head.vision_head.weight, head.vision_head.bias)]
        got = simpo_step_chunked_fp32(*w, torch.cat([hc, hr]), torch.cat([lc, lr]), T, L, chunk_rows=40, **hp)
        np.testing.assert_allclose(float(got["loss"]), float(ref["loss"]), rtol=1e-5)
        np.testing.assert_allclose(got["rejected_logps"].numpy(), ref["rejected_logps"].detach().numpy(), rtol=1e-5)
        for k in ("dx", "dW2", "dW1", "db2", "db1"):
            err = float((got[k] - ref[k]).norm() / ref[k].norm())
            assert err < 1e-5, (k, err)
