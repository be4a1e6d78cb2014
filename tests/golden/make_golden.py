"""Generate the golden fixtures in tests/golden/ by running the REAL reference code.

The reference package cannot be imported as-is in this container (transformers 5.x vs the pinned
4.38.2; pytorch_lightning / trl / attrdict / omegaconf / pyrootutils absent).  This script installs
inert stand-ins for exactly those third-party modules in ``sys.modules`` and then loads the reference
source files *unmodified* from /root/reference with importlib, so the arithmetic below is executed by
the reference's own lines:

  * ``vision_head``                         janus/models/modeling_vlm.py:36-51
  * ``model_name_to_cls``                   janus/models/modeling_vlm.py:54-70
  * ``JanusProTrainWrapper.get_batch_logps``   ospo/wrapper/train.py:375-396
  * ``JanusProTrainWrapper.simpo_loss``        ospo/wrapper/train.py:317-342
  * ``JanusProTrainWrapper.concatenated_forward`` / ``get_batch_loss_metrics``  train.py:345-372, 399-445
  * ``MlpProjector`` (gen_aligner)            janus/models/projector.py:27-86
  * ``JanusProImageGenWrapper.generate_image`` ospo/wrapper/image_generation.py:109-171
    (run with a stand-in backbone; ``torch.multinomial`` is intercepted to record the ``probs`` the
    reference hands to it)

It only runs where /root/reference exists (the build container).  The fixtures it writes are
committed; tests never need the reference at run time.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


# ------------------------------------------------------------------------------------------------
# stand-ins for third-party modules the reference imports but this container lacks
# ------------------------------------------------------------------------------------------------
class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs() -> None:
    _mod("attrdict", AttrDict=_AttrDict)

    class LightningModule(torch.nn.Module):
        """just enough of pl.LightningModule for the wrappers' arithmetic"""

        @property
        def device(self):
            return torch.device("cpu")

        def log(self, *a, **k):
            self.__dict__.setdefault("_logged", {})[a[0]] = a[1]

        def log_dict(self, d, *a, **k):
            self.__dict__.setdefault("_logged", {}).update(d)

    pl = _mod("pytorch_lightning", LightningModule=LightningModule, Trainer=object,
              seed_everything=lambda s, workers=False: torch.manual_seed(s))
    _mod("pytorch_lightning.strategies", DDPStrategy=object)
    _mod("pytorch_lightning.callbacks", ModelCheckpoint=object)
    _mod("pytorch_lightning.loggers", TensorBoardLogger=object)
    pl.strategies = sys.modules["pytorch_lightning.strategies"]

    def pad_to_length(tensor, length, pad_value, dim=-1):
        # trl.trainer.utils.pad_to_length (published behaviour): right-pad `dim` up to `length`
        if tensor.size(dim) >= length:
            return tensor
        pad_size = list(tensor.shape)
        pad_size[dim] = length - tensor.size(dim)
        return torch.cat([tensor, pad_value * torch.ones(*pad_size, dtype=tensor.dtype, device=tensor.device)],
                         dim=dim)

    _mod("trl")
    _mod("trl.trainer")
    _mod("trl.trainer.utils", pad_to_length=pad_to_length)
    _mod("pyrootutils", setup_root=lambda *a, **k: None)
    _mod("omegaconf", OmegaConf=object)
    # reference sub-packages whose own imports we do not need
    janus = _mod("janus")
    janus.__path__ = [str(REF / "janus")]
    jm = _mod("janus.models")
    jm.__path__ = [str(REF / "janus" / "models")]
    _mod("janus.models.clip_encoder", CLIPVisionTower=object)
    ospo = _mod("ospo")
    ospo.__path__ = [str(REF / "ospo")]
    ou = _mod("ospo.utils")
    ou.__path__ = [str(REF / "ospo" / "utils")]
    ow = _mod("ospo.wrapper")
    ow.__path__ = [str(REF / "ospo" / "wrapper")]
    _mod("ospo.utils.processor", get_conversation=None, get_sft_format=None)


class _PlainConfig:
    """stand-in for transformers.PretrainedConfig (v4 semantics: plain class, kwargs become attributes);
    transformers 5 turns every subclass into a dataclass, which the reference's v4-era config classes
    (mutable class-level defaults) do not survive.  The config classes are not on the golden path."""

    model_type = ""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            try:
                setattr(self, k, v)
            except AttributeError:
                pass


class _Registry:
    @staticmethod
    def register(*a, **k):
        return None


def load_modeling_vlm():
    """load janus/models/modeling_vlm.py with a v4-flavoured `transformers` facade for the duration"""
    saved = {k: sys.modules.get(k) for k in ("transformers", "transformers.configuration_utils")}
    _mod("transformers", AutoConfig=_Registry, AutoModelForCausalLM=_Registry, LlamaConfig=_PlainConfig,
         LlamaForCausalLM=torch.nn.Module, PreTrainedModel=torch.nn.Module)
    _mod("transformers.configuration_utils", PretrainedConfig=_PlainConfig)
    try:
        return load_ref("janus.models.modeling_vlm", "janus/models/modeling_vlm.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _bits(t: torch.Tensor) -> np.ndarray:
    assert t.dtype == torch.bfloat16
    return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16)


def load_ref(modname: str, relpath: str):
    spec = importlib.util.spec_from_file_location(modname, REF / relpath)
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------------------------
def main() -> None:
    install_stubs()
    load_ref("janus.models.projector", "janus/models/projector.py")
    mv = load_modeling_vlm()
    load_ref("ospo.constant", "ospo/constant.py")
    load_ref("ospo.utils.common", "ospo/utils/common.py")
    load_ref("ospo.utils.train", "ospo/utils/train.py")
    tr = load_ref("ospo.wrapper.train", "ospo/wrapper/train.py")
    ig = load_ref("ospo.wrapper.image_generation", "ospo/wrapper/image_generation.py")

    vision_head = mv.model_name_to_cls("vision_head")
    assert vision_head is mv.vision_head

    # ============================ SimPO golden =================================================
    H, E, V = 64, 96, 512
    B, T, L = 3, 24, 5
    g = torch.Generator().manual_seed(20251018)
    torch.manual_seed(20251018)
    head = vision_head(_AttrDict(n_embed=H, image_token_embed=E, image_token_size=V))
    # a bit more spread than default init so log-probs are not all ~ -log V
    with torch.no_grad():
        head.vision_head.weight.mul_(4.0)
    hidden_c = torch.randn(B, L + T, H, generator=g)
    hidden_r = torch.randn(B, L + T, H, generator=g)
    ids_c = torch.randint(0, V, (B, T), generator=g)
    ids_r = torch.randint(0, V, (B, T), generator=g)
    pad = torch.full((B, L), -100, dtype=torch.long)
    labels_c = torch.cat([pad, ids_c], dim=1)
    labels_r = torch.cat([pad, ids_r], dim=1)

    class _Backbone(torch.nn.Module):
        """stands in for language_model.model: returns the given embeds as last hidden state"""

        def forward(self, inputs_embeds=None, use_cache=False, past_key_values=None, **kw):
            return types.SimpleNamespace(hidden_states=[inputs_embeds], last_hidden_state=inputs_embeds)

    model = torch.nn.Module()
    model.gen_head = head
    model.language_model = torch.nn.Module()
    model.language_model.model = _Backbone()

    simpo = {}
    for tag, hp in {
        "sigmoid": dict(loss_type="sigmoid", beta=10.0, gamma_beta_ratio=0.5, label_smoothing=0.0, sft_weight=0.0),
        "sigmoid_smooth_sft": dict(loss_type="sigmoid", beta=2.0, gamma_beta_ratio=0.25, label_smoothing=0.1,
                                   sft_weight=0.5),
        "hinge": dict(loss_type="hinge", beta=1.0, gamma_beta_ratio=0.0, label_smoothing=0.0, sft_weight=0.0),
    }.items():
        w = tr.JanusProTrainWrapper.__new__(tr.JanusProTrainWrapper)
        torch.nn.Module.__init__(w)
        w.model = model
        for k, v in hp.items():
            setattr(w, k, v)
        w.label_pad_token_id = -100
        w.padding_value = 0
        xc = hidden_c.clone().requires_grad_(True)
        xr = hidden_r.clone().requires_grad_(True)
        batch = {
            "chosen_inputs_embeds": xc, "chosen_labels": labels_c,
            "chosen_attention_mask": torch.ones(B, L + T, dtype=torch.long),
            "rejected_inputs_embeds": xr, "rejected_labels": labels_r,
            "rejected_attention_mask": torch.ones(B, L + T, dtype=torch.long),
        }
        head.zero_grad()
        # pieces (reference methods, unmodified)
        cl, rl, clog, rlog, clab = w.concatenated_forward(batch)
        losses, crew, rrew = w.simpo_loss(cl, rl)
        # the whole step (reference get_batch_loss_metrics -> loss), then autograd
        loss = w.get_batch_loss_metrics(batch, "train")
        loss.backward()
        logged = {k: float(v) for k, v in w.__dict__["_logged"].items()}
        all_logits = torch.cat([clog, rlog], 0).detach()
        all_labels = torch.cat([labels_c, labels_r], 0)
        per_tok = torch.gather(all_logits[:, :-1].log_softmax(-1), 2,
                               all_labels[:, 1:].clamp(min=0).unsqueeze(2)).squeeze(2)
        simpo[tag] = dict(
            hp=hp, loss=float(loss.detach()), chosen_logps=cl.detach().numpy(), rejected_logps=rl.detach().numpy(),
            losses=losses.detach().numpy(), chosen_rewards=crew.numpy(), rejected_rewards=rrew.numpy(),
            per_token_logps=per_tok.numpy(), logged=logged,
            dx_chosen=xc.grad.numpy().copy(), dx_rejected=xr.grad.numpy().copy(),
            dW1=head.output_mlp_projector.weight.grad.numpy().copy(),
            db1=head.output_mlp_projector.bias.grad.numpy().copy(),
            dW2=head.vision_head.weight.grad.numpy().copy(), db2=head.vision_head.bias.grad.numpy().copy(),
        )
    flat = {}
    for tag, d in simpo.items():
        for k, v in d.items():
            if k == "hp":
                for hk, hv in v.items():
                    flat[f"{tag}/hp/{hk}"] = np.array(hv)
            elif k == "logged":
                for lk, lv in v.items():
                    flat[f"{tag}/logged/{lk}"] = np.array(lv)
            else:
                flat[f"{tag}/{k}"] = np.asarray(v)
    np.savez_compressed(
        OUT / "simpo_ref.npz",
        H=H, E=E, V=V, B=B, T=T, L=L,
        W1=head.output_mlp_projector.weight.detach().numpy(), b1=head.output_mlp_projector.bias.detach().numpy(),
        W2=head.vision_head.weight.detach().numpy(), b2=head.vision_head.bias.detach().numpy(),
        hidden_chosen=hidden_c.numpy(), hidden_rejected=hidden_r.numpy(),
        labels_chosen=labels_c.numpy(), labels_rejected=labels_r.numpy(),
        **flat,
    )
    print("wrote simpo_ref.npz:", {k: (v["loss"]) for k, v in simpo.items()})

    # ============================ CFG decode golden ===========================================
    # real generate_image loop (image_generation.py:109-171) with stand-in backbone / tokenizer /
    # VQ decoder; 3 decode steps, P = 2 prompts (4 CFG rows), bf16 head like utils/model.py:39.
    Hc, Ec, Vc, P, STEPS = 32, 32, 16384, 2, 3
    torch.manual_seed(7)
    head_c = vision_head(_AttrDict(n_embed=Hc, image_token_embed=Ec, image_token_size=Vc))
    with torch.no_grad():
        head_c.vision_head.weight.mul_(6.0)
    head_c = head_c.to(torch.bfloat16)
    gh = torch.Generator().manual_seed(99)
    hidden_steps = torch.randn(STEPS, 2 * P, Hc, generator=gh).to(torch.bfloat16)
    recorded = {"probs": [], "logits": []}

    class _LM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.step = 0

        def forward(self, inputs_embeds=None, attention_mask=None, use_cache=True, past_key_values=None):
            h = hidden_steps[self.step]
            self.step += 1
            last = torch.zeros(inputs_embeds.shape[0], inputs_embeds.shape[1], Hc, dtype=torch.bfloat16)
            last[:, -1, :] = h
            return types.SimpleNamespace(last_hidden_state=last, past_key_values=object())

    class _GenHeadRec(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x):
            out = self.inner(x)
            recorded["logits"].append(out.detach().clone())
            return out

    mdl = torch.nn.Module()
    mdl.gen_head = _GenHeadRec(head_c)
    mdl.language_model = torch.nn.Module()
    mdl.language_model.model = _LM()
    emb = torch.nn.Embedding(32, Hc).to(torch.bfloat16)
    mdl.language_model.get_input_embeddings = lambda: emb
    mdl.prepare_gen_img_embeds = lambda ids: torch.zeros(ids.shape[0], Hc, dtype=torch.bfloat16)
    mdl.gen_vision_model = types.SimpleNamespace(
        decode_code=lambda toks, shape: torch.zeros(shape[0], 3, 384, 384))

    gw = ig.JanusProImageGenWrapper.__new__(ig.JanusProImageGenWrapper)
    torch.nn.Module.__init__(gw)
    gw.model = mdl
    gw.processor = types.SimpleNamespace(tokenizer=types.SimpleNamespace(encode=lambda p: [1, 5, 6, 7, 2]), pad_id=0)
    gw.cfg_weight = 5.0
    gw.temperature = 1.0
    ig.set_seed = lambda seed: torch.manual_seed(seed if seed is not None else 0)

    real_multinomial = torch.multinomial

    def rec_multinomial(probs, num_samples=1, **kw):
        recorded["probs"].append(probs.detach().clone())
        return real_multinomial(probs, num_samples, **kw)

    # the reference hard-codes .cuda() for the token buffer (image_generation.py:147); run it on CPU
    _zeros = torch.zeros
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.multinomial = rec_multinomial
    # The reference runs this loop on CUDA under Lightning's bf16 autocast, whose documented policy
    # promotes softmax to float32 (CPU autocast does not).  Emulate exactly that one promotion.
    real_softmax = torch.softmax
    torch.softmax = lambda x, dim=-1, **kw: real_softmax(x.float(), dim=dim, **kw)
    ig.Image = types.SimpleNamespace(fromarray=lambda a: types.SimpleNamespace(save=lambda p: None))
    try:
        gw.generate_image(["a", "b"], ["/tmp/_g0.png", "/tmp/_g1.png"], seed=3, image_token_num_per_image=STEPS)
    finally:
        torch.multinomial = real_multinomial
        torch.softmax = real_softmax
    probs = torch.stack(recorded["probs"])           # [STEPS, P, V] fp32
    logits = torch.stack(recorded["logits"])         # [STEPS, 2P, V] bf16
    assert probs.shape == (STEPS, P, Vc) and logits.shape == (STEPS, 2 * P, Vc)
    np.savez_compressed(
        OUT / "cfg_ref.npz",
        H=Hc, E=Ec, V=Vc, P=P, STEPS=STEPS, cfg_weight=5.0, temperature=1.0,
        # bf16 tensors are stored as their uint16 bit patterns (bf16 = upper half of fp32)
        W1_bf16=_bits(head_c.output_mlp_projector.weight), b1_bf16=_bits(head_c.output_mlp_projector.bias),
        W2_bf16=_bits(head_c.vision_head.weight), b2_bf16=_bits(head_c.vision_head.bias),
        hidden_bf16=_bits(hidden_steps), logits_bf16=_bits(logits),
        probs=probs.numpy(), greedy=probs.argmax(-1).numpy(),
    )
    print("wrote cfg_ref.npz: probs", tuple(probs.shape), "max p", float(probs.max()))


def aligner_golden() -> None:
    """real MlpProjector (projector.py) + nn.Embedding exactly as MultiModalityCausalLM.prepare_gen_img_embeds wires
    them (modeling_vlm.py:204-216, 263-264), bf16 like the generation path (utils/model.py:39)"""
    proj = sys.modules["janus.models.projector"]
    torch.manual_seed(11)
    D, CB = 128, 512
    gen_aligner = proj.MlpProjector(_AttrDict(projector_type="mlp_gelu", depth=2, input_dim=8, n_embed=D))
    gen_embed = torch.nn.Embedding(CB, 8)
    gen_aligner = gen_aligner.to(torch.bfloat16)
    gen_embed = gen_embed.to(torch.bfloat16)
    ids = torch.randint(0, CB, (20,), generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        out = gen_aligner(gen_embed(ids))          # == prepare_gen_img_embeds(ids)
    np.savez_compressed(
        OUT / "aligner_ref.npz", D=D, CB=CB, ids=ids.numpy(),
        gen_embed_bf16=_bits(gen_embed.weight), wa_bf16=_bits(gen_aligner.layers[0].weight),
        ba_bf16=_bits(gen_aligner.layers[0].bias), wb_bf16=_bits(gen_aligner.layers[2].weight),
        bb_bf16=_bits(gen_aligner.layers[2].bias), out_bf16=_bits(out))
    print("wrote aligner_ref.npz:", tuple(out.shape))


def adamw_golden() -> None:
    """next row N3.  The reference's optimizer is PyTorch's own (ospo/wrapper/train.py:108-115 torch.optim.AdamW,
    Lightning gradient_clip_val -> torch.nn.utils.clip_grad_norm_, ospo/utils/train.py:30,50): run exactly those on
    a tiny head-shaped parameter set for three steps, with the step-5 hyper-parameters (configs/step5.yaml:37-43)
    and with a weight-decay / no-clip variant."""
    H, E, V = 24, 40, 64
    out = {}
    for tag, kw, max_norm, gscale in (("a", dict(lr=4e-5, betas=(0.9, 0.95), weight_decay=0.0, eps=1e-8), 1.0, 3.0),
                                      ("b", dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8), 0.0, 0.1)):
        g = torch.Generator().manual_seed(21)
        shapes = [(V, E), (E, H), (V,), (E,)]                      # W2, W1, b2, b1: the flat-buffer order
        params = [torch.nn.Parameter(torch.randn(*s, generator=g) * 0.05) for s in shapes]
        opt = torch.optim.AdamW(params, foreach=False, fused=False, **kw)
        out[f"{tag}_p0"] = np.concatenate([p.detach().numpy().ravel() for p in params])
        for step in range(3):
            grads = [torch.randn(*s, generator=g) * gscale * (0.2 if step == 2 else 1.0) for s in shapes]
            for p_, g_ in zip(params, grads):
                p_.grad = g_.clone()
            out[f"{tag}_g{step}"] = np.concatenate([x.numpy().ravel() for x in grads])
            if max_norm > 0:
                tn = torch.nn.utils.clip_grad_norm_(params, max_norm)
                out[f"{tag}_norm{step}"] = np.float32(tn.item())
            opt.step()
            out[f"{tag}_p{step + 1}"] = np.concatenate([p.detach().numpy().ravel() for p in params])
            st = [opt.state[p_] for p_ in params]
            out[f"{tag}_m{step + 1}"] = np.concatenate([s_["exp_avg"].numpy().ravel() for s_ in st])
            out[f"{tag}_v{step + 1}"] = np.concatenate([s_["exp_avg_sq"].numpy().ravel() for s_ in st])
        out[f"{tag}_hyper"] = np.array([kw["lr"], kw["betas"][0], kw["betas"][1], kw["eps"], kw["weight_decay"], max_norm],
                                       dtype=np.float64)
    np.savez_compressed(OUT / "adamw_ref.npz", H=H, E=E, V=V, torch_version=torch.__version__, **out)
    print("wrote adamw_ref.npz")


if __name__ == "__main__":
    if "--only-adamw" in sys.argv:
        adamw_golden()
    else:
        main()
        aligner_golden()
        adamw_golden()
