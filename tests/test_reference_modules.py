"""The drop-in behind the reference's OWN modules: `source.layers.Attention`, `Transformer` and the full
`TransformingSRT` built from the reference's YAMLs run on CUDA with and without `gta_b200.gta.install()`, forward and
backward, fp32 and under the reference's bf16 autocast (source/trainer.py:106), and are compared with the reference's
eager fp32 result on the same weights and inputs.  The reference is imported from /root/reference (build container) or
baseline/_ref (GPU box; baseline/install_ref.py) — unmodified either way.

Budgets: forward-only fp32 1e-3, bf16 / autocast / training 1e-2, both relative to max(1, |ref|_max) of the compared
tensor (a module output passes through the to_out projection, so its scale is not O(1))."""
import copy

import pytest
import torch

from baseline import ref_loader, srt_synth

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not available (run baseline/install_ref.py)")
RUNS = ["msn/GTA/gta_so3", "clevrtr/GTA/gta"]


def _attention_modules(model):
    enc = model.encoder.transformer.layers[0][0].fn
    dec = model.decoder.allocation_transformer.transformer.layers[0][0].fn
    return enc, dec


def _rel(a, b):
    return float((a.float() - b.float()).abs().max()) / max(1.0, float(b.float().abs().max()))


@pytest.fixture()
def fast():
    from gta_b200 import gta as fast
    yield fast
    fast.uninstall()


# ------------------------------------------------------------------------------------------------ CPU (no GPU needed)
@needs_ref
def test_tau_is_read_from_the_attn_fn_closure():
    """`tau` is a closure variable of AttnFn.forward (source/layers.py:195-211), never a keyword argument."""
    from gta_b200 import gta as fast
    ref = ref_loader.load()
    args = {"method": {"name": "gta", "args": {"f_dims": {"se3": 16, "so2": 16}, "so2": 4}}}
    a = ref.layers.Attention(64, heads=2, dim_head=32, attn_args=args)
    assert fast._closure_tau(a.attn_fn) == (1.0, True)
    b = ref.layers.Attention(64, heads=2, dim_head=32, attn_args=dict(args, softmax="adjustable"))
    tau, found = fast._closure_tau(b.attn_fn)
    assert found and tau is b.attend.tau and isinstance(tau, torch.nn.Parameter)
    e = ref.layers.Attention(64, heads=2, dim_head=32, attn_args={"method": {"name": "gta", "args": dict(
        args["method"]["args"], euclid_sim=True)}, "softmax": "adjustable"})
    assert fast._closure_tau(e.attn_fn)[0] is e.attend.tau          # EuclidAttnFn closes over the same variable

    class Plain:
        scale = 1.0
    assert fast._closure_tau(Plain()) == (1.0, False)


@needs_ref
def test_install_rebinds_rep_builders_and_cpu_calls_reach_the_reference(fast):
    """install() swaps the attention function and both pre_compute_reps methods; with CPU tensors everything is handed to
    the reference's own code (and produces its exact result)."""
    ref = ref_loader.load()
    run = "clevrtr/GTA/gta"
    torch.manual_seed(0)
    model, _ = srt_synth.build_model(ref, run, "cpu", dropout=0.0)
    b = srt_synth.make_batch(run, 1, "cpu")
    ex = lambda: {k: b[k] for k in ("input_transforms", "target_transforms", "input_coord", "target_coord")}
    with torch.no_grad():
        want, _ = model(b["input_images"], b["input_camera_pos"], b["input_rays"], b["target_camera_pos"], b["target_rays"], ex())
    orig_enc = ref.encoder.ImprovedSRTEncoder.pre_compute_reps
    fast.install()
    assert ref.encoder.ImprovedSRTEncoder.pre_compute_reps is not orig_enc
    assert ref.layers.multihead_geometric_transform_attention is fast.multihead_geometric_transform_attention
    with torch.no_grad(), pytest.warns(UserWarning):
        fast._warned.clear()
        got, _ = model(b["input_images"], b["input_camera_pos"], b["input_rays"], b["target_camera_pos"], b["target_rays"], ex())
    assert torch.equal(got, want)
    fast.uninstall()
    assert ref.encoder.ImprovedSRTEncoder.pre_compute_reps is orig_enc


# ------------------------------------------------------------------------------------------------ GPU
def _module_case(ref, run, which, B, seed=0):
    torch.manual_seed(seed)
    model, cfg = srt_synth.build_model(ref, run, "cuda", dropout=0.0)
    enc, dec = _attention_modules(model)
    batch = srt_synth.make_batch(run, B, "cuda", seed=seed)
    extras = {k: batch[k] for k in ("input_transforms", "target_transforms", "input_coord", "target_coord")}
    Ni, Nt, H, W, h, w, P = srt_synth.shapes_for(run)
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    if which == "enc":
        mod, x, z = enc, torch.randn(B, Ni * h * w, enc.to_qkv.in_features, device="cuda", generator=g), None
    else:
        mod = dec
        x = torch.randn(B, Nt * P, dec.to_q.in_features, device="cuda", generator=g)
        z = torch.randn(B, Ni * h * w, dec.to_kv.in_features, device="cuda", generator=g)
    return model, mod, x, z, extras, cfg


def _reps(model, which, extras):
    ex = dict(extras)
    model.encoder.pre_compute_reps(model.encoder.attn_args, ex)
    if which == "dec":
        model.decoder.pre_compute_reps(model.decoder.attn_args, ex)
    return ex


def _fwd_bwd(mod, x, z, ex, autocast, grad=True):
    x = x.clone().requires_grad_(grad)
    z = None if z is None else z.clone().requires_grad_(grad)
    mod.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast), torch.set_grad_enabled(grad):
        out = mod(x, z=z, extras=ex)
    res = {"out": out.detach().float()}
    if grad:
        gen = torch.Generator(device="cuda").manual_seed(5)
        out.float().backward(torch.randn(out.shape, device="cuda", generator=gen))
        res["dx"] = x.grad.float()
        if z is not None:
            res["dz"] = z.grad.float()
        res["dtc"] = mod.trans_coeff.grad.float().clone()
        w = mod.to_qkv.weight if z is None else mod.to_kv.weight
        res["dW"] = w.grad.float().clone()
    return res


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("which", ["enc", "dec"])
@pytest.mark.parametrize("run", RUNS)
def test_reference_attention_module_with_install(run, which, fast):
    """source.layers.Attention.forward (layers.py:292,388-430) of the encoder (self-attention) and of the decoder
    (cross-attention), reference eager fp32 vs the drop-in: fp32 forward-only, fp32 training, bf16 autocast training."""
    ref = ref_loader.load()
    fast.uninstall()
    model, mod, x, z, extras, cfg = _module_case(ref, run, which, B=2)
    with torch.no_grad():
        mod.trans_coeff.fill_(0.05)
    want = _fwd_bwd(mod, x, z, _reps(model, which, extras), autocast=False)
    fast.install()
    assert ref.layers.multihead_geometric_transform_attention is fast.multihead_geometric_transform_attention
    ex = _reps(model, which, extras)
    assert ex[fast._PACK_KEY][0] == "built" and "so2rep_q" not in ex, "the installed pre_compute_reps must build the tables on the device"
    report = {}
    # forward only, fp32: split-precision kernel
    got = _fwd_bwd(mod, x, z, ex, autocast=False, grad=False)
    report["fp32_fwd"] = _rel(got["out"], want["out"])
    assert report["fp32_fwd"] < 1e-3, report
    # training, fp32 inputs (bf16 tensor-core math both ways) and under the reference's bf16 autocast
    for name, ac in (("fp32_train", False), ("autocast_train", True)):
        got = _fwd_bwd(mod, x, z, _reps(model, which, extras), autocast=ac)
        for k in want:
            report[name + "_" + k] = _rel(got[k], want[k])
            budget = 1e-2 if k != "dtc" else 5e-2       # d(trans_coeff): four strongly cancelling parts
            if ac and k != "out":
                budget *= 3          # the projections around the op run in bf16 under autocast as well (both arms of the comparison would)
            assert report[name + "_" + k] < budget, (k, report)
    print("max-abs / max(1,|ref|):", {k: "%.2e" % v for k, v in report.items()})


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("run", RUNS)
def test_full_model_train_step_with_install(run, fast):
    """TransformingSRT from the YAML (encoder 5 + decoder 2 GTA layers): loss, prediction and gradients of one training
    step (source/trainer.py:69-125) with the drop-in installed vs the reference's eager path, same weights and batch,
    in the precision the config trains in (MSN: bf16 autocast, CLEVR: fp32)."""
    ref = ref_loader.load()
    fast.uninstall()
    torch.manual_seed(1)
    model, cfg = srt_synth.build_model(ref, run, "cuda", dropout=0.0)
    mixed = bool(cfg["training"].get("mixed_prec", False))
    batch = srt_synth.make_batch(run, 2, "cuda", seed=3)

    def step(mp):
        model.zero_grad(set_to_none=True)
        loss, pred = srt_synth.loss_fn(model, batch, mp)
        loss.backward()
        grads = {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}
        return float(loss), pred.detach().float(), grads

    loss_ref32, pred_ref32, g_ref32 = step(False)               # ground truth: the reference in fp32
    loss_ref, pred_ref, g_ref = step(mixed)                     # the reference in the config's training precision
    fast.install()
    loss_got, pred_got, g_got = step(mixed)
    fast.uninstall()
    e_got, e_ref = _rel(pred_got, pred_ref32), _rel(pred_ref, pred_ref32)
    print("pred err vs fp32 reference: drop-in %.2e, reference in training precision %.2e; loss %.6f / %.6f / %.6f"
          % (e_got, e_ref, loss_got, loss_ref, loss_ref32))
    assert e_got < 1e-2
    assert abs(loss_got - loss_ref32) < 1e-2 * max(1e-3, abs(loss_ref32)) + 1e-4
    assert set(g_got) == set(g_ref32)
    cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
    worst = min((cos(g_got[n], g_ref32[n]), n) for n in g_got if g_ref32[n].norm() > 0)
    worst_ref = min((cos(g_ref[n], g_ref32[n]), n) for n in g_ref if g_ref32[n].norm() > 0)
    print("worst gradient cosine vs fp32 reference: drop-in %.5f (%s), reference in training precision %.5f (%s)"
          % (worst + worst_ref))
    tcs = [n for n in g_got if n.endswith("trans_coeff")]
    assert len(tcs) == 7 and all(torch.isfinite(g_got[n]).all() for n in tcs)
    assert worst[0] > min(0.98, worst_ref[0] - 0.02), (worst, worst_ref)


@pytest.mark.gpu
@needs_ref
def test_adjustable_softmax_temperature(fast):
    """`softmax: adjustable` (layers.py:195-200): the drop-in reads tau from the closure; when tau needs a gradient it
    comes from the identity d(loss)/d(tau) = -(1/tau) sum(q * dq) on the library's dq."""
    ref = ref_loader.load()
    fast.uninstall()
    torch.manual_seed(2)
    args = {"method": {"name": "gta", "args": {"f_dims": {"se3": 32, "so2": 32}, "so2": 8, "max_freq_h": 1, "max_freq_w": 1}},
            "softmax": "adjustable"}
    att = ref.layers.Attention(384, heads=6, dim_head=64, attn_args=args).cuda()
    with torch.no_grad():
        att.attend.tau.fill_(1.7)
    enc = ref.encoder.ImprovedSRTEncoder.__new__(ref.encoder.ImprovedSRTEncoder)
    batch = srt_synth.make_batch("clevrtr/GTA/gta", 2, "cuda", seed=4)
    base = {k: batch[k] for k in ("input_transforms", "input_coord")}
    x = torch.randn(2, 600, 384, device="cuda")

    def reps():
        ex = dict(base)
        ref.encoder.ImprovedSRTEncoder.pre_compute_reps(enc, args["method"]["args"], ex)
        return ex
    with torch.no_grad():
        want = att(x, extras=reps())
        att.attend.tau.fill_(1.0)
        other = att(x, extras=reps())
        att.attend.tau.fill_(1.7)
    assert _rel(other, want) > 1e-2                       # the temperature matters on these inputs
    fast.install()
    with torch.no_grad():
        got = att(x, extras=reps())
    assert _rel(got, want) < 1e-3
    w = torch.randn_like(want)
    (att(x, extras=reps()) * w).sum().backward()
    got_tau, got_tc = att.attend.tau.grad.clone(), att.trans_coeff.grad.clone()
    fast.uninstall()
    att.zero_grad()
    (att(x, extras=reps()) * w).sum().backward()
    ref_tau, ref_tc = att.attend.tau.grad, att.trans_coeff.grad
    print("d tau: drop-in %.5f reference %.5f;  d trans_coeff: %.5f / %.5f" % (float(got_tau), float(ref_tau), float(got_tc), float(ref_tc)))
    assert abs(float(got_tau) - float(ref_tau)) <= 2e-2 * abs(float(ref_tau)) + 1e-3


@pytest.mark.gpu
@needs_ref
def test_return_attmap_is_honoured_per_call(fast):
    """Transformer(return_last_attmap=True) asks its last Attention for the map (layers.py:441-444,478-480; heads == 1):
    the drop-in returns it for exactly that call and None otherwise."""
    ref = ref_loader.load()
    fast.uninstall()
    torch.manual_seed(3)
    args = {"method": {"name": "gta", "args": {"f_dims": {"se3": 48, "so3": 24, "so2": 24}, "so2": 6, "so3": 2,
                                               "max_freq_h": 1, "max_freq_w": 1}}}
    tr = ref.layers.Transformer(96, depth=2, heads=1, dim_head=96, mlp_dim=192, selfatt=True, return_last_attmap=True,
                                attn_args=args).cuda()
    enc = ref.encoder.ImprovedSRTEncoder.__new__(ref.encoder.ImprovedSRTEncoder)
    batch = srt_synth.make_batch("msn/GTA/gta_so3", 2, "cuda", seed=6)
    base = {"input_transforms": batch["input_transforms"][:, :2], "input_coord": batch["input_coord"][:, :2, :64]}
    x = torch.randn(2, 128, 96, device="cuda")

    def reps():
        ex = dict(base)
        ref.encoder.ImprovedSRTEncoder.pre_compute_reps(enc, args["method"]["args"], ex)
        return ex
    with torch.no_grad():
        want, want_map = tr(x, None, reps())
        fast.install()
        got, got_map = tr(x, None, reps())
    assert got_map is not None and got_map.shape == want_map.shape == (2, 1, 128, 128)
    assert float((got_map - want_map).abs().max()) < 1e-3
    assert _rel(got, want) < 1e-3
    # direct call without a requesting caller: no map
    att = tr.layers[0][0].fn
    with torch.no_grad():
        q = torch.randn(2, 1, 128, 96, device="cuda")
        o, m = fast.multihead_geometric_transform_attention(q, q, q, att.attn_fn, args["method"]["args"]["f_dims"], reps(),
                                                            trans_coeff=att.trans_coeff)
    assert m is None
