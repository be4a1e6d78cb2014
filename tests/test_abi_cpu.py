"""CPU-side checks: the C-ABI library loads and exports every symbol include/ospo_head.h declares, the
ctypes mirror matches the header, argument validation works without a GPU, and the host-side row
selection (the label shift / mask of get_batch_logps) matches the oracle."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from ospo_b200 import _abi
from ospo_b200.head import FusedGenHead, _rows_from_labels

ROOT = Path(__file__).resolve().parent.parent


def _header_functions():
    text = (ROOT / "include" / "ospo_head.h").read_text()
    return set(re.findall(r"OSPO_API\s+[\w\s\*]+?\b(ospo_head_\w+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(_abi.lib_path()))
    declared = _header_functions()
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_header_constants_match_python_mirror():
    text = (ROOT / "include" / "ospo_head.h").read_text()
    consts = {k: int(v) for k, v in re.findall(r"#define\s+(OSPO_\w+)\s+(-?\d+)\b", text)}
    assert consts["OSPO_SC_LOSS"] == _abi.SC_LOSS and consts["OSPO_SC_COUNT"] == _abi.SC_COUNT
    assert consts["OSPO_SC_LOGITS_REJECTED"] == _abi.SC_LOGITS_REJECTED
    assert consts["OSPO_LOSS_HINGE"] == _abi.LOSS_HINGE and consts["OSPO_MERGE_FP32"] == _abi.MERGE_FP32
    assert consts["OSPO_K_COUNT"] == len(_abi.KERNEL_NAMES)


def test_strerror_and_workspace_query_need_no_gpu():
    assert _abi.strerror(0) == "ok"
    assert "sm_100" in _abi.strerror(-5)
    n = _abi.workspace_bytes(73728, 4096, 4096, 16384, 128)
    # LSE partials (128 column sub-tiles x rows x 12 B) + dpre and the row-weighted activations (rows x E x 2 B each)
    # dominate
    lo = 73728 * 128 * 12 + 2 * 73728 * 4096 * 2
    assert lo <= n < 1.1 * lo
    with pytest.raises(_abi.OspoHeadError):
        _abi.workspace_bytes(0, 8, 8, 8)


def test_calls_fail_loudly_without_a_b200():
    """no CPU fallback: on a machine without an sm_100 device every compute entry point reports an error"""
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less machines")
    lib = _abi.load()
    args = _abi.HeadArgs()
    rc = lib.ospo_head_logits(C.byref(args), None)
    assert rc != 0
    head = FusedGenHead(type("P", (), dict(n_embed=16, image_token_embed=16, image_token_size=64)))
    with pytest.raises(Exception):
        with torch.no_grad():
            head(torch.zeros(2, 16))


def test_state_dict_is_checkpoint_compatible_with_reference_head():
    from oracle.head_oracle import VisionHead

    ref = VisionHead(32, 48, 64)
    head = FusedGenHead(type("P", (), dict(n_embed=32, image_token_embed=48, image_token_size=64)))
    assert list(head.state_dict().keys()) == list(ref.state_dict().keys())
    head.load_state_dict(ref.state_dict(), strict=True)
    adopted = FusedGenHead.from_reference(ref)
    assert adopted.vision_head.weight is ref.vision_head.weight
    ref.vision_head.weight.requires_grad_(False)
    assert not adopted.vision_head.weight.requires_grad


@pytest.mark.parametrize("use_span", [False, True])
def test_row_selection_matches_get_batch_logps_mask(use_span):
    from oracle import head_oracle as O

    S, Lt, T, H, V = 4, 3, 6, 8, 32
    g = torch.Generator().manual_seed(0)
    hidden = torch.randn(S, Lt + T, H, generator=g)
    labels = torch.cat([torch.full((S, Lt), -100), torch.randint(0, V, (S, T), generator=g)], 1)
    span = (Lt - 1, Lt - 1 + T) if use_span else None
    x_rows, targets, seq_off, seg = _rows_from_labels(hidden, labels, -100, span)
    assert seg == (0, 0)   # CPU / fp32 / T % 64 != 0: the gather path
    lab = labels[:, 1:]
    mask = lab != -100
    assert torch.equal(x_rows, hidden[:, :-1][mask])
    assert torch.equal(targets, lab[mask])
    assert seq_off.tolist() == [0, T, 2 * T, 3 * T, 4 * T]
    # and those are exactly the positions the oracle's get_batch_logps averages over
    head = O.make_head(H, 8, V, seed=1)
    _, per_tok, m = O.get_batch_logps(head(hidden), labels, return_per_token=True)
    assert torch.equal(m, mask)


def test_total_grad_norm_matches_reference_loop():
    """N2: the sync-free grad-norm equals the reference's per-parameter .item() loop (train.py:464-469)"""
    import torch

    from ospo_b200.patch import total_grad_norm

    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(3))]
    for p in ps[:2]:
        p.grad = torch.randn_like(p)
    total = 0.0
    for p in ps:
        if p.grad is not None:
            total += p.grad.detach().data.norm(2).item() ** 2
    got = total_grad_norm(ps)
    assert abs(float(got) - total ** 0.5) < 1e-5
    assert total_grad_norm([torch.nn.Parameter(torch.zeros(2))]) == 0.0


def test_token_handoff_matches_reference_pixels(tmp_path):
    """N4: ids -> decode_code -> uint8 NHWC -> PNG, byte-identical to the reference's numpy path
    (image_generation.py:174-191) on a stand-in decoder"""
    import numpy as np
    import torch
    from PIL import Image

    from ospo_b200.handoff import save_images, tokens_to_uint8

    P, img, patch = 3, 48, 16
    side = img // patch
    g = torch.Generator().manual_seed(3)
    tokens = torch.randint(0, 16384, (P, side * side), generator=g).to(torch.int64)
    table = torch.randn(16384, 3, generator=g) * 0.9
    seen = {}

    def decode_code(code, shape):                       # vq_model.py:505-508 signature
        seen["dtype"], seen["shape"] = code.dtype, shape
        base = table[code.long()].view(P, side, side, 3).permute(0, 3, 1, 2)
        return torch.nn.functional.interpolate(base, size=(img, img), mode="bilinear").to(torch.bfloat16)

    got = tokens_to_uint8(decode_code, tokens, img_size=img, patch_size=patch)
    assert seen["dtype"] == torch.int32 and seen["shape"] == [P, 8, side, side]
    # the reference's lines, literally
    dec = decode_code(tokens.to(dtype=torch.int), shape=[P, 8, side, side])
    dec = dec.to(torch.float32).cpu().numpy().transpose(0, 2, 3, 1)
    dec = np.clip((dec + 1) / 2 * 255, 0, 255)
    visual_img = np.zeros((P, img, img, 3), dtype=np.uint8)
    visual_img[:, :, :] = dec
    assert got.dtype == np.uint8 and np.array_equal(got, visual_img)
    paths = [str(tmp_path / f"img_{i:02d}.png") for i in range(P)]
    assert save_images(got, paths) == paths
    assert np.array_equal(np.asarray(Image.open(paths[1])), visual_img[1])


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """the drop-in boundary is a C ABI: include/ospo_head.h compiles as C99 (and C++11) without CUDA or torch headers,
    and a C program that only includes it links against libospo_head.so and can call a function that needs no GPU"""
    import shutil
    import subprocess
    from pathlib import Path

    from ospo_b200 import _abi

    root = Path(__file__).resolve().parent.parent
    hdr = root / "include" / "ospo_head.h"
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no C compiler")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", str(hdr)], check=True)
    subprocess.run([shutil.which("g++") or gcc, "-std=c++11", "-fsyntax-only", "-x", "c++", str(hdr)], check=True)
    lib = _abi.lib_path()
    if not lib.exists():
        import pytest
        pytest.skip("library not built")
    src = tmp_path / "t.c"
    src.write_text(
        '#include <stdio.h>\n#include "ospo_head.h"\n'
        "int main(void) {\n"
        "  ospo_head_shape s = {73728, 4096, 4096, 16384, 128};\n"
        "  size_t need = 0;\n"
        "  int rc = ospo_head_workspace_bytes(&s, &need);\n"
        '  printf("%d %zu %s\\n", rc, need, ospo_head_strerror(-5));\n'
        "  return rc;\n}\n")
    exe = tmp_path / "t"
    subprocess.run([gcc, "-std=c99", "-I", str(root / "include"), str(src), "-o", str(exe), str(lib),
                    f"-Wl,-rpath,{lib.parent}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(maxsplit=2)
    assert out[0] == "0" and int(out[1]) == _abi.workspace_bytes(73728, 4096, 4096, 16384, 128)
    assert "sm_100" in out[2]


# ---------------------------------------------------------------------------------------------------
# next row N2: batched VQ encode in the train-wrapper patch (ospo/wrapper/train.py:219-279)
# ---------------------------------------------------------------------------------------------------
def _serial_preprocess_batch(self, batch):
    """literal restatement of JanusProTrainWrapper.preprocess_batch (ospo/wrapper/train.py:219-279): per-sample text
    embedding with zero padding, one batch-1 VQ encode per image (:246-261), two prepare_gen_img_embeds calls"""
    import torch

    batch_size = len(batch[0])
    item_ids, text_tokens, chosen_image_tensors, rejected_image_tensors = batch
    embs = [self.model.language_model.get_input_embeddings()(t) for t in text_tokens]
    max_seq_len = max(x.shape[1] for x in embs)
    padded = torch.zeros(batch_size, max_seq_len, embs[0].size(-1), dtype=self.model.dtype, device=self.device)
    labels = torch.full((batch_size, max_seq_len), -100, dtype=torch.long, device=self.device)
    for i, e in enumerate(embs):
        padded[i, :e.shape[1], :] = e
    cl, rl = [], []
    for c, r in zip(chosen_image_tensors, rejected_image_tensors):
        cl.append(self.model.gen_vision_model.encode(c.to(self.device))[2][2])
        rl.append(self.model.gen_vision_model.encode(r.to(self.device))[2][2])
    ct, rt = torch.stack(cl, 0), torch.stack(rl, 0)
    ce, re_ = self.model.prepare_gen_img_embeds(ct), self.model.prepare_gen_img_embeds(rt)
    out = {"item_ids": item_ids}
    out["chosen_inputs_embeds"] = torch.cat([padded, ce], dim=1)
    out["chosen_attention_mask"] = torch.ones(out["chosen_inputs_embeds"].shape[:2], dtype=torch.long)
    out["chosen_labels"] = torch.cat([labels, ct], dim=1)
    out["rejected_inputs_embeds"] = torch.cat([padded, re_], dim=1)
    out["rejected_attention_mask"] = torch.ones(out["rejected_inputs_embeds"].shape[:2], dtype=torch.long)
    out["rejected_labels"] = torch.cat([labels, rt], dim=1)
    return out


def _reference_vq_model():
    """the reference's own VQ tokenizer (janus/models/vq_model.py, torch only) when the reference tree is present,
    else a stand-in with the same encode() contract: (quant, losses, (perplexity, min_encodings, indices))"""
    import importlib.util
    import os

    import torch

    path = "/root/reference/janus/models/vq_model.py"
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("ref_vq_model", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(3)
        return mod.VQ_16().eval(), "reference VQ_16"

    class StandIn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, 2, 1), torch.nn.SiLU(), torch.nn.Conv2d(16, 8, 3, 8, 1))
            self.codebook = torch.nn.Embedding(512, 8)

        def encode(self, x):
            z = self.enc(x).permute(0, 2, 3, 1).reshape(-1, 8)
            idx = torch.cdist(z, self.codebook.weight).argmin(-1)
            return None, None, (None, None, idx)

    torch.manual_seed(3)
    return StandIn().eval(), "stand-in"


def test_batched_preprocess_batch_matches_the_serial_reference_loop():
    """ids, embeddings and labels of the batched VQ encode equal the reference's serial batch-1 loop bit for bit (CPU,
    fp32: the convolutions are evaluated per sample either way); the token cache returns the same ids without
    calling the encoder again"""
    import types

    import torch

    from ospo_b200.patch import batched_preprocess_batch

    vq, which = _reference_vq_model()
    torch.manual_seed(4)
    model = torch.nn.Module()
    model.language_model = torch.nn.Module()
    emb = torch.nn.Embedding(100, 32)
    model.language_model.get_input_embeddings = lambda: emb
    model.gen_vision_model = vq
    gen_embed = torch.nn.Embedding(16384, 32)
    model.prepare_gen_img_embeds = lambda ids: gen_embed(ids)
    model.dtype = torch.float32
    w = types.SimpleNamespace(model=model, device=torch.device("cpu"))
    B, S = 3, 64                                          # 64 x 64 images -> 4 x 4 = 16 tokens with the 16x encoder
    g = torch.Generator().manual_seed(5)
    batch = ([f"item{i}" for i in range(B)],
             [torch.randint(0, 100, (1, n), generator=g) for n in (5, 9, 2)],
             [torch.rand(1, 3, S, S, generator=g) * 2 - 1 for _ in range(B)],
             [torch.rand(1, 3, S, S, generator=g) * 2 - 1 for _ in range(B)])
    with torch.no_grad():
        ref = _serial_preprocess_batch(w, batch)
        cache = {}
        got = batched_preprocess_batch(w, batch, cache)
        calls = []
        orig = vq.encode
        vq.encode = lambda x: calls.append(x.shape) or orig(x)
        again = batched_preprocess_batch(w, batch, cache)
    assert not calls, f"{which}: cached items must not be re-encoded"
    for out in (got, again):
        assert out["item_ids"] == ref["item_ids"]
        for k in ("chosen_labels", "rejected_labels", "chosen_attention_mask", "rejected_attention_mask"):
            assert torch.equal(out[k], ref[k]), (which, k)
        for k in ("chosen_inputs_embeds", "rejected_inputs_embeds"):
            assert torch.equal(out[k], ref[k]), (which, k)
    assert ref["chosen_labels"].shape[1] == 9 + (S // 16) ** 2


def test_packed_exp_rounding_step_is_not_contracted_in_the_sass():
    """The samplers' n = rint(t log2 e) is  (rn(t * log2 e) + 1.5 * 2^23) - 1.5 * 2^23  on the packed fp32 pipe
    (cfg_math.cuh); oracle/cfg_sample.c rounds the product before it rounds the sum.  ptxas contracts
    mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 when it can, which rounds once and moves n at ties.  In every kernel of
    the library each packed '+ 12582912' (FADD2, or FFMA2 with the multiplier 1 of f2_add_nofuse) must therefore
    have its own packed multiply by log2 e in front of it."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_abi.lib_path())], capture_output=True, text=True).stdout
    per_fn, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = [0, 0]
            continue
        if fn is None:
            continue
        if re.search(r"\bFMUL2\b.*1\.44269502", line):
            per_fn[fn][0] += 1
        elif re.search(r"\b(FADD2|FFMA2)\b.*, 12582912 ;", line):
            per_fn[fn][1] += 1
    users = {k: v for k, v in per_fn.items() if v[1]}
    assert users, "no kernel with the packed rounding step found: did the SASS mnemonics change?"
    assert any("cfg_merge_sample_kernel" in k for k in users)
    for k, (n_mul, n_magic) in users.items():
        assert n_mul == n_magic, (k, n_mul, n_magic)
