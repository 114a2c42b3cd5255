"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference modules from /root/reference.

Only usable in the build container (the reference tree does not exist on the GPU box). Used by
oracle/gen_golden.py to generate tests/golden/*.pt and by the CPU tests that pin the restatements
in oracle/restate_*.py against the real reference. Nothing under vitron_b200/ imports this.

The shims follow SURVEY.md Appendix C: stub the absent third-party packages (peft, decord,
pytorchvideo, xformers, open_clip, fairscale, rotary_embedding_torch), pre-register empty package
modules so the reference's `__init__` chains (which import MPT) are skipped, and paper over
transformers 4.31 -> 5.x API drift. No file under /root/reference is modified.
"""
import importlib
import importlib.util
import os
import sys
import types

REF = os.environ.get("VITRON_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "vitron"))


def _stub(name, **attrs):
    if name in sys.modules:
        m = sys.modules[name]
    else:
        m = types.ModuleType(name)
        m.__path__ = []  # behave as a package so that sub-imports resolve
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def load_file(modname, relpath):
    """Import a self-contained reference file by path."""
    path = os.path.join(REF, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_vitron_ready = False


def setup_vitron():
    """Make `vitron.model.language_model.llava_llama` & co importable."""
    global _vitron_ready
    if _vitron_ready:
        return
    import torch  # noqa: F401
    import transformers
    from transformers.models.clip import modeling_clip

    if REF not in sys.path:
        sys.path.insert(0, REF)
    _pkg("vitron", os.path.join(REF, "vitron"))
    _pkg("vitron.model", os.path.join(REF, "vitron/model"))
    _pkg("vitron.model.language_model", os.path.join(REF, "vitron/model/language_model"))

    class _LoraConfig:  # peft stub
        def __init__(self, *a, **k):
            pass

    _stub("peft", LoraConfig=_LoraConfig, get_peft_model=lambda m, c: m, PeftModel=object)
    if not hasattr(modeling_clip, "_expand_mask"):
        modeling_clip._expand_mask = lambda mask, dtype, tgt_len=None: mask
    transformers.AutoConfig.register = staticmethod(lambda *a, **k: None)
    transformers.AutoModelForCausalLM.register = staticmethod(lambda *a, **k: None)

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, k):
            return _Any()

    for name in ("decord", "pytorchvideo", "pytorchvideo.data", "pytorchvideo.data.encoded_video",
                 "pytorchvideo.transforms", "torchvision.transforms._transforms_video"):
        m = _stub(name)
        m.__getattr__ = lambda k, _A=_Any: _A()  # module-level __getattr__ (PEP 562)
    sys.modules["decord"].VideoReader = _Any
    sys.modules["decord"].cpu = lambda *a, **k: None

    try:
        from transformers.cache_utils import DynamicCache
        if not hasattr(DynamicCache, "_vb_getitem"):
            def _getitem(self, i):
                layer = self.layers[i]
                return (layer.keys, layer.values)
            DynamicCache.__getitem__ = _getitem
            DynamicCache._vb_getitem = True
    except Exception:  # pragma: no cover
        pass
    _vitron_ready = True


def vitron_classes():
    """Returns a namespace with the reference classes of the Vicuna / ViT / adapter path."""
    setup_vitron()
    ns = types.SimpleNamespace()
    ll = importlib.import_module("vitron.model.language_model.llava_llama")
    ns.LlavaLlamaForCausalLM = ll.LlavaLlamaForCausalLM
    ns.LlavaConfig = ll.LlavaConfig
    ns.llava_arch = importlib.import_module("vitron.model.llava_arch")
    lb = importlib.import_module("vitron.model.multimodal_encoder.languagebind")
    ns.LanguageBindImageTower = lb.LanguageBindImageTower
    ns.LanguageBindVideoTower = lb.LanguageBindVideoTower
    mi = importlib.import_module("vitron.model.multimodal_encoder.languagebind.image.modeling_image")
    mv = importlib.import_module("vitron.model.multimodal_encoder.languagebind.video.modeling_video")
    ns.ImageVisionTransformer = mi.CLIPVisionTransformer
    ns.VideoVisionTransformer = mv.CLIPVisionTransformer
    ci = importlib.import_module("vitron.model.multimodal_encoder.languagebind.image.configuration_image")
    cv = importlib.import_module("vitron.model.multimodal_encoder.languagebind.video.configuration_video")
    ns.ImageVisionConfig = ci.CLIPVisionConfig
    ns.VideoVisionConfig = cv.CLIPVisionConfig
    pj = importlib.import_module("vitron.model.multimodal_projector.builder")
    ns.build_vision_projector = pj.build_vision_projector
    rg = importlib.import_module("vitron.model.region_extractor.builder")
    ns.build_region_extractor = rg.build_region_extractor
    rl = importlib.import_module("vitron.model.region_extractor.layer")
    ns.RegionExtractor = rl.RegionExtractor
    return ns


def build_reference_vitron(llm_cfg, vit_cfg, with_video=False, num_frames=8, hidden_act="gelu", seed=0):
    """Construct the reference LlavaLlamaForCausalLM with attached towers / projector / region
    extractor (random init, fp32, eval) without touching from_pretrained.

    llm_cfg: dict(hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, vocab_size)
    vit_cfg: dict(hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, image_size, patch_size)
    """
    import torch
    import torch.nn as nn
    ns = vitron_classes()
    torch.manual_seed(seed)
    cfg = ns.LlavaConfig(**llm_cfg, rms_norm_eps=1e-5, max_position_embeddings=4096, pad_token_id=0,
                         bos_token_id=1, eos_token_id=2, attn_implementation="eager")
    cfg.pretraining_tp = 1
    cfg.mm_projector_type = "mlp2x_gelu"
    cfg.mm_hidden_size = vit_cfg["hidden_size"]
    cfg.tokenizer_padding_side = "right"
    cfg.tokenizer_model_max_length = 4096
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = ns.LlavaLlamaForCausalLM(cfg)

    def tower(cls, vt_cls, vcfg_cls, attr, **extra):
        vcfg = vcfg_cls(hidden_size=vit_cfg["hidden_size"], intermediate_size=vit_cfg["intermediate_size"],
                        num_hidden_layers=vit_cfg["num_hidden_layers"],
                        num_attention_heads=vit_cfg["num_attention_heads"], image_size=vit_cfg["image_size"],
                        patch_size=vit_cfg["patch_size"], hidden_act=hidden_act, layer_norm_eps=1e-5,
                        lora_r=0, attn_implementation="eager", **extra)
        vit = vt_cls(vcfg)
        t = cls.__new__(cls)
        nn.Module.__init__(t)
        t.is_loaded = True
        t.select_layer = -2
        t.select_feature = "patch"
        setattr(t, attr, vit)
        return t

    model.model.image_tower = tower(ns.LanguageBindImageTower, ns.ImageVisionTransformer, ns.ImageVisionConfig,
                                    "image_tower", add_time_attn=False, num_frames=1, force_patch_dropout=0.0)
    if with_video:
        model.model.video_tower = tower(ns.LanguageBindVideoTower, ns.VideoVisionTransformer, ns.VideoVisionConfig,
                                        "video_tower", add_time_attn=True, num_frames=num_frames,
                                        force_patch_dropout=0.0)
    model.model.mm_projector = ns.build_vision_projector(cfg)
    rx = ns.RegionExtractor(in_dim=vit_cfg["hidden_size"], out_dim=llm_cfg["hidden_size"],
                            patch_size=vit_cfg["patch_size"], image_size=vit_cfg["image_size"])
    model.model.region_extractor = rx
    model.eval()
    return model


# ------------------------------------------------------------------ i2vgen-xl UNet
_i2v_ready = False


def setup_i2vgen():
    global _i2v_ready
    if _i2v_ready:
        return
    import torch
    import torch.nn.functional as F
    base = os.path.join(REF, "modules/i2vgen-xl")
    if base not in sys.path:
        sys.path.insert(0, base)
    _pkg("tools", os.path.join(base, "tools"))
    _pkg("tools.modules", os.path.join(base, "tools/modules"))
    _pkg("tools.modules.unet", os.path.join(base, "tools/modules/unet"))

    def mea(q, k, v, attn_bias=None, op=None):
        return F.scaled_dot_product_attention(q, k, v)

    xo = _stub("xformers.ops", memory_efficient_attention=mea, LowerTriangularMask=object)
    _stub("xformers", ops=xo)
    _stub("open_clip")

    class _Rot:
        def __init__(self, *a, **k):
            pass

    _stub("rotary_embedding_torch", RotaryEmbedding=_Rot)
    _stub("fairscale")
    _stub("fairscale.nn")
    _stub("fairscale.nn.checkpoint", checkpoint_wrapper=lambda m, *a, **k: m)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # unet_i2vgen.py:283 hard-codes .cuda()
    _i2v_ready = True


def i2vgen_unet_class():
    setup_i2vgen()
    mod = importlib.import_module("tools.modules.unet.unet_i2vgen")
    return mod.UNetSD_I2VGen


def i2vgen_autoencoder_class():
    """Unmodified tools/modules/autoencoder.py::AutoencoderKL (its `utils.registry_class` import resolves inside
    modules/i2vgen-xl, which setup_i2vgen puts on sys.path)."""
    setup_i2vgen()
    mod = importlib.import_module("tools.modules.autoencoder")
    return mod.AutoencoderKL


def i2vgen_ddim_class():
    setup_i2vgen()
    base = os.path.join(REF, "modules/i2vgen-xl")
    _pkg("tools.modules.diffusions", os.path.join(base, "tools/modules/diffusions"))
    mod = importlib.import_module("tools.modules.diffusions.diffusion_ddim")
    return mod.DiffusionDDIM


def gligen_attention():
    return load_file("ref_gligen_attention", "modules/GLIGEN/demo/gligen/ldm/modules/attention.py")


def gligen_unet_class():
    """Unmodified gligen/ldm/modules/diffusionmodules/openaimodel.py::UNetModel. Its imports are absolute from the
    Vitron repo root (`modules.GLIGEN.demo.gligen.ldm...`), so the reference root goes on sys.path."""
    chain = ["modules", "modules.GLIGEN", "modules.GLIGEN.demo", "modules.GLIGEN.demo.gligen", "modules.GLIGEN.demo.gligen.ldm"]
    for name in chain:  # empty package modules: skips gligen/__init__.py and ldm/__init__.py (which import its trainer / evaluator stack)
        if name not in sys.modules:
            _pkg(name, os.path.join(REF, *name.split(".")))
    return importlib.import_module("modules.GLIGEN.demo.gligen.ldm.modules.diffusionmodules.openaimodel").UNetModel


def gligen_autoencoder_class():
    """Unmodified gligen/ldm/models/autoencoder.py::AutoencoderKL (encode = posterior.sample() * scale_factor,
    decode(z) = decoder(post_quant_conv(z / scale_factor)))."""
    gligen_unet_class()
    name = "modules.GLIGEN.demo.gligen.ldm.models"
    if name not in sys.modules:
        _pkg(name, os.path.join(REF, *name.split(".")))
    return importlib.import_module("modules.GLIGEN.demo.gligen.ldm.models.autoencoder").AutoencoderKL


def gligen_plms_classes():
    """Unmodified (PLMSSampler, DDPM) of gligen/ldm/models/diffusion/{plms,ddpm}.py (same package stubs as the UNet)."""
    gligen_unet_class()
    for name in ("modules.GLIGEN.demo.gligen.ldm.models", "modules.GLIGEN.demo.gligen.ldm.models.diffusion"):
        if name not in sys.modules:
            _pkg(name, os.path.join(REF, *name.split(".")))
    plms = importlib.import_module("modules.GLIGEN.demo.gligen.ldm.models.diffusion.plms")
    ddpm = importlib.import_module("modules.GLIGEN.demo.gligen.ldm.models.diffusion.ddpm")
    return plms.PLMSSampler, ddpm.DDPM


def seem_pieces():
    """The importable SEEM pieces (torch/einops only)."""
    ns = types.SimpleNamespace()
    base = "modules/SEEM/demo_code/xdecoder"
    ns.attn = load_file("ref_seem_attn", base + "/utils/attn.py")
    return ns


_seem_ready = False


def setup_seem():
    """Make the *unmodified* SEEM classes importable: `xdecoder.body.decoder.seem`
    (MultiScaleMaskedTransformerDecoder), `xdecoder.body.encoder.transformer_encoder_fpn`
    (TransformerEncoderPixelDecoder) and `xdecoder.backbone.focal` (FocalNet).

    Absent third-party packages are replaced by the minimal layer definitions the reference uses
    from them (stated here because they are OUR statement of third-party code, pinned versions in
    modules/SEEM/requirements: detectron2@afe9eb9, timm 0.9.16, fvcore):
      * detectron2.layers.Conv2d  = nn.Conv2d + optional `norm` module + optional `activation`
        callable applied in that order (detectron2/layers/wrappers.py `Conv2d.forward`);
      * detectron2.layers.get_norm("GN", c) = nn.GroupNorm(32, c); "" -> None;
      * detectron2.layers.ShapeSpec = namedtuple(channels, height, width, stride);
      * fvcore.nn.weight_init.c2_xavier_fill / c2_msra_fill: init only (weights are overwritten);
      * timm.models.layers: trunc_normal_ (init only), to_2tuple, DropPath (identity in eval);
      * omegaconf.DictConfig: only used in an isinstance() check of `configurable`.
    The xdecoder package `__init__` chains (which pull the language encoder, registry side effects
    and detectron2 structures) are skipped by pre-registering empty package modules.
    """
    global _seem_ready
    if _seem_ready:
        return
    import collections

    import torch
    from torch import nn

    class Conv2d(nn.Conv2d):
        def __init__(self, *args, **kwargs):
            norm = kwargs.pop("norm", None)
            activation = kwargs.pop("activation", None)
            super().__init__(*args, **kwargs)
            self.norm = norm
            self.activation = activation

        def forward(self, x):
            x = nn.functional.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
            if self.norm is not None:
                x = self.norm(x)
            if self.activation is not None:
                x = self.activation(x)
            return x

    def get_norm(norm, out_channels):
        if norm is None or norm == "":
            return None
        assert norm == "GN", norm
        return nn.GroupNorm(32, out_channels)

    class ShapeSpec(collections.namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
        def __new__(cls, channels=None, height=None, width=None, stride=None):
            return super().__new__(cls, channels, height, width, stride)

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training, "oracle shim: DropPath is identity (eval only)"
            return x

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean, std, a, b)

    noop = lambda *a, **k: None
    _stub("detectron2")
    _stub("detectron2.layers", Conv2d=Conv2d, DeformConv=None, ShapeSpec=ShapeSpec, get_norm=get_norm,
          cat=torch.cat, shapes_to_tensor=None)
    _stub("detectron2.utils")
    _stub("detectron2.utils.file_io", PathManager=None)
    _stub("detectron2.modeling", BACKBONE_REGISTRY=None, Backbone=nn.Module, ShapeSpec=ShapeSpec)
    _stub("detectron2.structures", BitMasks=None, Boxes=None)
    _stub("fvcore")
    _stub("fvcore.nn")
    _stub("fvcore.nn.weight_init", c2_xavier_fill=noop, c2_msra_fill=noop)
    sys.modules["fvcore.nn"].weight_init = sys.modules["fvcore.nn.weight_init"]
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", trunc_normal_=trunc_normal_, to_2tuple=to_2tuple, DropPath=DropPath)
    _stub("omegaconf", DictConfig=type("DictConfig", (), {}))

    base = os.path.join(REF, "modules/SEEM/demo_code/xdecoder")
    rel = "modules/SEEM/demo_code/xdecoder"
    _pkg("xdecoder", base)
    cfgmod = load_file("xdecoder.utils.config", rel + "/utils/config.py")
    u = _pkg("xdecoder.utils", os.path.join(base, "utils"))
    u.configurable = cfgmod.configurable
    pe = load_file("xdecoder.modules.position_encoding", rel + "/modules/position_encoding.py")
    pf = load_file("xdecoder.modules.point_features", rel + "/modules/point_features.py")
    m = _pkg("xdecoder.modules", os.path.join(base, "modules"))
    m.PositionEmbeddingSine = pe.PositionEmbeddingSine
    m.point_features = pf
    _pkg("xdecoder.body", os.path.join(base, "body"))
    _pkg("xdecoder.body.decoder", os.path.join(base, "body/decoder"))
    _pkg("xdecoder.body.encoder", os.path.join(base, "body/encoder"))
    _pkg("xdecoder.backbone", os.path.join(base, "backbone"))
    _seem_ready = True


def seem_classes():
    """(TransformerEncoderPixelDecoder, MultiScaleMaskedTransformerDecoder, ShapeSpec) — unmodified."""
    setup_seem()
    enc = importlib.import_module("xdecoder.body.encoder.transformer_encoder_fpn")
    dec = importlib.import_module("xdecoder.body.decoder.seem")
    return enc.TransformerEncoderPixelDecoder, dec.MultiScaleMaskedTransformerDecoder, sys.modules["detectron2.layers"].ShapeSpec


def seem_focalnet_class():
    setup_seem()
    return importlib.import_module("xdecoder.backbone.focal").FocalNet
