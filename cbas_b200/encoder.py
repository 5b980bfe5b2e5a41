"""DinoEncoder - host-side mirror of the reference's encoder object over the sm_100a C ABI.

Mirrors `backend/cbas.py:650-677` (class DinoEncoder): same constructor arguments, a `.device` attribute,
and `__call__(x[B,S,H,W] float in [0,1]) -> [B,S,D]`.  The ViT arithmetic the reference delegates to
`transformers.AutoModel` (modeling_dinov3_vit.py) runs in libcbas_b200.so instead; PyTorch is used only for
device memory, streams and the one-off weight re-layout below.

Two generalisations over the reference, both called out in SURVEY.md:
  * the embedding width is `config.hidden_size`, not the hard-coded 768 of cbas.py:677;
  * `encode_u8` takes the decoder's uint8 RGB frames directly (green extraction, 1/255 and the 3x channel
    replication of cbas.py:431,674 are folded into the patch-embedding weights), and a PROCESSOR mode
    reproduces HF's DINOv3ViTImageProcessor (rescale -> antialiased bilinear resize -> ImageNet normalise).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

PRE_REFERENCE = 0
PRE_PROCESSOR = 1

# hub configs of facebook/dinov3-vit{s,b,l}16-pretrain-lvd1689m (recalled; weights are gated)
ARCHITECTURES: Dict[str, Dict[str, int]] = {
    "vits16": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536),
    "vitb16": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072),
    "vitl16": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
    # DINOv2-with-registers (CBAS's default encoder, cbas.py:1030-1033): 14-pixel patches, learned absolute position
    # embedding trained on a 37x37 grid (518 px), no RoPE, layer_norm_eps 1e-6
    "dinov2reg-s14": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536,
                          patch_size=14, layer_norm_eps=1e-6, family="dinov2_with_registers", pos_grid=37),
    "dinov2reg-b14": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                          patch_size=14, layer_norm_eps=1e-6, family="dinov2_with_registers", pos_grid=37),
}


@dataclass
class ViTConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    num_register_tokens: int = 4
    patch_size: int = 16
    layer_norm_eps: float = 1e-5
    rope_theta: float = 100.0
    family: str = "dinov3_vit"   # or "dinov2_with_registers"
    pos_grid: int = 0            # dinov2: side of the trained position-embedding grid

    @classmethod
    def from_hf(cls, cfg) -> "ViTConfig":
        if getattr(cfg, "model_type", "dinov3_vit") == "dinov2_with_registers":
            if getattr(cfg, "use_swiglu_ffn", False):
                raise ValueError("SwiGLU DINOv2 variants (giant) are not supported by the B200 encoder")
            if getattr(cfg, "hidden_act", "gelu") != "gelu" or not getattr(cfg, "qkv_bias", True):
                raise ValueError("only hidden_act='gelu' with qkv_bias is supported")
            if int(cfg.patch_size) not in (14, 16) or int(cfg.num_channels) != 3:
                raise ValueError("only 14- or 16-pixel patches on 3-channel input are supported")
            size = cfg.image_size if isinstance(cfg.image_size, int) else cfg.image_size[0]
            return cls(cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                       int(cfg.hidden_size * cfg.mlp_ratio), cfg.num_register_tokens, int(cfg.patch_size),
                       float(cfg.layer_norm_eps), 0.0, "dinov2_with_registers", size // int(cfg.patch_size))
        if getattr(cfg, "use_gated_mlp", False):
            raise ValueError("gated-MLP DINOv3 variants are not supported by the B200 encoder")
        if getattr(cfg, "hidden_act", "gelu") != "gelu":
            raise ValueError("only hidden_act='gelu' is supported")
        if int(cfg.patch_size) != 16 or int(cfg.num_channels) != 3:
            raise ValueError("only 16-pixel patches on 3-channel input are supported")
        return cls(cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size,
                   cfg.num_register_tokens, 16, float(cfg.layer_norm_eps), float(cfg.rope_theta))


def rope_tables(n_h: int, n_w: int, head_dim: int = 64, theta: float = 100.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin of the DINOv3 axial RoPE for an n_h x n_w patch grid, first half of the head dim only
    (the reference table is `angles.tile(2)`, modeling_dinov3_vit.py:168-200, so cos[i+32] == cos[i]).
    Coordinates are patch centres in [-1,1] (modeling_dinov3_vit.py:95-121); all arithmetic in fp32."""
    ch = torch.arange(0.5, n_h, dtype=torch.float32) / n_h
    cw = torch.arange(0.5, n_w, dtype=torch.float32) / n_w
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    inv_freq = 1 / theta ** torch.arange(0, 1, 4 / head_dim, dtype=torch.float32)
    angles = 2 * math.pi * coords[:, :, None] * inv_freq[None, None, :]
    angles = angles.flatten(1, 2)  # [Np, head_dim/2] = [y*f0..f15, x*f0..f15]
    return torch.cos(angles).contiguous(), torch.sin(angles).contiguous()


def aa_bilinear_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """First source index and normalised weights of torch's antialiased bilinear resize along one axis
    (ATen _upsample_bilinear2d_aa weight computation, align_corners=False), fp32 like the float kernel.
    Returns (xmin[out] int32, w[out, taps] float32)."""
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    support = f32(1.0) * scale if scale >= 1.0 else f32(1.0)
    invscale = f32(1.0) / scale if scale >= 1.0 else f32(1.0)
    taps = int(math.ceil(float(support))) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, taps), np.float32)
    for i in range(out_size):
        center = scale * f32(i + 0.5)
        lo = max(int(center - support + f32(0.5)), 0)
        hi = min(int(center + support + f32(0.5)), in_size)
        ws = np.zeros(taps, np.float32)
        total = f32(0.0)
        for j in range(hi - lo):
            x = abs((f32(j + lo) - center + f32(0.5)) * invscale)
            v = f32(1.0) - x if x < 1.0 else f32(0.0)
            ws[j] = v
            total += v
        if total != 0:
            ws = (ws / total).astype(np.float32)
        xmin[i] = lo
        w[i] = ws
    return xmin, w


def synthetic_state_dict(cfg: ViTConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random weights with the key layout and init scale of transformers' DINOv3ViTModel
    (_init_weights, modeling_dinov3_vit.py:470-487: trunc_normal(std=0.02) weights, zero biases, LayerScale 1).
    Used by bench.py and smoke(): the gated hub weights are unavailable offline."""
    g = torch.Generator().manual_seed(seed)
    D, I = cfg.hidden_size, cfg.intermediate_size

    def tn(*shape):
        t = torch.empty(*shape)
        nn.init.trunc_normal_(t, mean=0.0, std=0.02, generator=g)
        return t

    sd = {
        "embeddings.cls_token": tn(1, 1, D),
        "embeddings.register_tokens": tn(1, cfg.num_register_tokens, D),
        "embeddings.patch_embeddings.weight": tn(D, 3, cfg.patch_size, cfg.patch_size),
        "embeddings.patch_embeddings.bias": torch.zeros(D),
        "norm.weight": torch.ones(D), "norm.bias": torch.zeros(D),
    }
    for i in range(cfg.num_hidden_layers):
        p = f"model.layer.{i}."
        sd.update({
            p + "norm1.weight": torch.ones(D), p + "norm1.bias": torch.zeros(D),
            p + "norm2.weight": torch.ones(D), p + "norm2.bias": torch.zeros(D),
            p + "attention.q_proj.weight": tn(D, D), p + "attention.q_proj.bias": torch.zeros(D),
            p + "attention.k_proj.weight": tn(D, D),
            **({p + "attention.k_proj.bias": torch.zeros(D)} if cfg.family == "dinov2_with_registers" else {}),
            p + "attention.v_proj.weight": tn(D, D), p + "attention.v_proj.bias": torch.zeros(D),
            p + "attention.o_proj.weight": tn(D, D), p + "attention.o_proj.bias": torch.zeros(D),
            p + "layer_scale1.lambda1": torch.ones(D), p + "layer_scale2.lambda1": torch.ones(D),
            p + "mlp.up_proj.weight": tn(I, D), p + "mlp.up_proj.bias": torch.zeros(I),
            p + "mlp.down_proj.weight": tn(D, I), p + "mlp.down_proj.bias": torch.zeros(D),
        })
    if cfg.family == "dinov2_with_registers":
        sd["embeddings.position_embeddings"] = tn(1, 1 + cfg.pos_grid * cfg.pos_grid, D)
    return sd


def normalize_state_dict(sd: Dict[str, torch.Tensor], family: str) -> Dict[str, torch.Tensor]:
    """Map a transformers Dinov2WithRegistersModel state dict onto the DINOv3 key names the packer below reads
    (modeling_dinov2_with_registers.py: query/key/value, output.dense, mlp.fc1/fc2, layernorm)."""
    if family != "dinov2_with_registers" or "norm.weight" in sd:
        return sd
    out = {}
    for k, v in sd.items():
        k = k.replace("embeddings.patch_embeddings.projection.", "embeddings.patch_embeddings.")
        k = k.replace("encoder.layer.", "model.layer.")
        for a, b in (("attention.attention.query.", "attention.q_proj."), ("attention.attention.key.", "attention.k_proj."),
                     ("attention.attention.value.", "attention.v_proj."), ("attention.output.dense.", "attention.o_proj."),
                     ("mlp.fc1.", "mlp.up_proj."), ("mlp.fc2.", "mlp.down_proj.")):
            k = k.replace(a, b)
        if k.startswith("layernorm."):
            k = "norm." + k[len("layernorm."):]
        out[k] = v
    return out


def interpolate_pos_embed(pos: torch.Tensor, n_h: int, n_w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Dinov2WithRegistersEmbeddings.interpolate_pos_encoding (modeling_dinov2_with_registers.py:93-145):
    pos [1, 1+G*G, D] -> (class position [D], patch positions [n_h*n_w, D]); bicubic, antialiased, fp32, and no
    resampling at all when the grid already matches."""
    pos = pos.float()
    cls_pos, patch_pos = pos[0, 0], pos[0, 1:]
    G = int(round(patch_pos.shape[0] ** 0.5))
    if G * G == n_h * n_w and n_h == n_w:
        return cls_pos, patch_pos
    D = pos.shape[-1]
    grid = patch_pos.reshape(1, G, G, D).permute(0, 3, 1, 2)
    grid = nn.functional.interpolate(grid, size=(n_h, n_w), mode="bicubic", align_corners=False, antialias=True)
    return cls_pos, grid.permute(0, 2, 3, 1).reshape(n_h * n_w, D)


def fold_layernorm(w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm(gamma, beta) followed by Linear(w, b), rewritten for the fused kernels (csrc/gemm_tcgen05.cuh):
    LN(h) W^T + b = rstd * (h - mean) W'^T + c2 with W' = W * gamma.  Returns (W' as bf16, c1, c2) where
    c1[n] = sum_k W'[n,k] is taken over the bf16-ROUNDED W' (it must cancel the tensor-core sum exactly when a row of
    h is constant) and c2 = W beta + b; sums in float64."""
    w64, g64 = w.double(), gamma.double()
    w_folded = (w64 * g64[None, :]).to(torch.bfloat16)
    c1 = w_folded.double().sum(dim=1).float()
    c2 = (w64 @ beta.double() + b.double()).float()
    return w_folded, c1, c2


def normalize_dinov3_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """transformers 5.x names the DINOv3 blocks `model.layer.{i}.*`; the 4.5x releases the reference's
    requirements.txt also admits (`transformers>=4.53.3`) use `layer.{i}.*`.  Accept both (and an optional leading
    `model.` from a wrapping module) and return the 5.x layout the packer reads."""
    if any(k.startswith("model.layer.") for k in sd):
        return sd
    out = {}
    for k, v in sd.items():
        if k.startswith("model.embeddings.") or k.startswith("model.norm.") or k.startswith("model.model.layer."):
            k = k[len("model."):]
        if k.startswith("layer."):
            k = "model." + k
        out[k] = v
    return out


class _NativeEncoder:
    """One libcbas_b200 encoder handle for a fixed input geometry; owns the device weight tensors."""

    def __init__(self, cfg: ViTConfig, sd: Dict[str, torch.Tensor], device: torch.device, mode: int,
                 in_hw: Tuple[int, int], side: int, max_frames: int):
        self.lib = _lib.lib()
        self.cfg, self.device, self.mode, self.in_hw, self.side = cfg, device, mode, in_hw, side
        self.max_frames = max_frames
        D = cfg.hidden_size
        self._keep = []  # device tensors whose pointers the library holds

        def dev(t: torch.Tensor, dtype) -> int:
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        f32, bf16 = torch.float32, torch.bfloat16
        sd = normalize_dinov3_keys(normalize_state_dict(sd, cfg.family))
        P = cfg.patch_size
        wp = sd["embeddings.patch_embeddings.weight"].float()
        if mode == PRE_REFERENCE:
            # three identical channels of G/255 (cbas.py:431,674)  ==  one channel against sum_c W / 255
            w_patch = (wp.sum(dim=1) / 255.0).reshape(D, P * P)
        else:
            w_patch = wp.reshape(D, 3 * P * P)
        kp = (w_patch.shape[1] + 63) // 64 * 64  # GEMM K granularity; the library zero-pads the patch matrix alike
        if kp != w_patch.shape[1]:
            w_patch = torch.cat([w_patch, torch.zeros(D, kp - w_patch.shape[1])], dim=1)
        prefix = torch.cat([sd["embeddings.cls_token"].reshape(1, D),
                            sd["embeddings.register_tokens"].reshape(-1, D)], dim=0).float().clone()
        n_side = side // P  # floor, like the stride-P convolution
        pos_patch = None
        if cfg.family == "dinov2_with_registers":
            cls_pos, pos_patch = interpolate_pos_embed(sd["embeddings.position_embeddings"], n_side, n_side)
            prefix[0] += cls_pos  # the CLS token gets its position before the registers are spliced in
        else:
            cos, sin = rope_tables(n_side, n_side, D // cfg.num_attention_heads, cfg.rope_theta)

        I = cfg.intermediate_size

        def opt(key: str, n: int) -> torch.Tensor:
            # the HF configs allow query_bias / value_bias / proj_bias / mlp_bias = False: an absent bias is zero
            t = sd.get(key)
            return t.float() if t is not None else torch.zeros(n)

        layers = (_lib.LayerWeights * cfg.num_hidden_layers)()
        for i in range(cfg.num_hidden_layers):
            p = f"model.layer.{i}."
            q_w, k_w, v_w = (sd[p + f"attention.{n}_proj.weight"].float() for n in "qkv")
            qkv_b = torch.cat([opt(p + f"attention.{n}_proj.bias", D) for n in "qkv"])  # key_bias=False in DINOv3
            l1 = sd[p + "layer_scale1.lambda1"].float()
            l2 = sd[p + "layer_scale2.lambda1"].float()
            lw = layers[i]
            # norm1 / norm2 (modeling_dinov3_vit.py:433,445) folded into the projection behind them
            wq, c1, c2 = fold_layernorm(torch.cat([q_w, k_w, v_w], dim=0), qkv_b,
                                        sd[p + "norm1.weight"], sd[p + "norm1.bias"])
            lw.w_qkv, lw.c1_qkv, lw.b_qkv = dev(wq, bf16), dev(c1, f32), dev(c2, f32)
            wu, c1, c2 = fold_layernorm(sd[p + "mlp.up_proj.weight"], opt(p + "mlp.up_proj.bias", I),
                                        sd[p + "norm2.weight"], sd[p + "norm2.bias"])
            lw.w_up, lw.c1_up, lw.b_up = dev(wu, bf16), dev(c1, f32), dev(c2, f32)
            # LayerScale (modeling_dinov3_vit.py:337-343) folded into the projection that feeds it
            lw.w_o = dev(sd[p + "attention.o_proj.weight"].float() * l1[:, None], bf16)
            lw.b_o = dev(opt(p + "attention.o_proj.bias", D) * l1, f32)
            lw.w_down = dev(sd[p + "mlp.down_proj.weight"].float() * l2[:, None], bf16)
            lw.b_down = dev(opt(p + "mlp.down_proj.bias", D) * l2, f32)
        self._layers = layers

        w = _lib.EncoderWeights()
        w.w_patch, w.b_patch = dev(w_patch, bf16), dev(sd["embeddings.patch_embeddings.bias"], f32)
        w.prefix = dev(prefix, f32)
        if pos_patch is None:
            w.rope_cos, w.rope_sin = dev(cos, f32), dev(sin, f32)
        else:
            w.pos_embed = dev(pos_patch, f32)
        w.lnf_g, w.lnf_b = dev(sd["norm.weight"], f32), dev(sd["norm.bias"], f32)
        w.layers = C.cast(layers, C.POINTER(_lib.LayerWeights))
        taps_y = taps_x = 0
        if mode == PRE_PROCESSOR:
            ymin, wy = aa_bilinear_taps(in_hw[0], side)
            xmin, wx = aa_bilinear_taps(in_hw[1], side)
            taps_y, taps_x = wy.shape[1], wx.shape[1]
            w.rs_ymin, w.rs_wy = dev(torch.from_numpy(ymin), torch.int32), dev(torch.from_numpy(wy), f32)
            w.rs_xmin, w.rs_wx = dev(torch.from_numpy(xmin), torch.int32), dev(torch.from_numpy(wx), f32)
        c = _lib.EncoderCfg(D, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size,
                            1 + cfg.num_register_tokens, mode, in_hw[0], in_hw[1], side, max_frames,
                            cfg.layer_norm_eps, taps_y, taps_x, P)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.cbas_b200_encoder_create(C.byref(c), C.byref(w), C.byref(handle)), "encoder_create")
        self.handle = handle
        self.tokens = n_side * n_side + 1 + cfg.num_register_tokens

    def set_option(self, option: int, value: int) -> None:
        _lib.check(self.lib.cbas_b200_encoder_set_option(self.handle, option, value), "encoder_set_option")

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.cbas_b200_encoder_destroy(h)
            except Exception:
                pass
            self.handle = None


class DinoEncoder(nn.Module):
    """Drop-in for `cbas.DinoEncoder` (cbas.py:650-677) backed by libcbas_b200.so.

    model_identifier: a Hugging Face id (loaded with transformers like the reference, cbas.py:657) or
        "synthetic:vit{s,b,l}16[@seed]" for random-init weights of that architecture (offline benchmarking).
    preprocess: "reference" (cbas.py:431,674: green/255 x3 at native resolution) or "processor"
        (HF DINOv3ViTImageProcessor: resize to `image_size`, ImageNet-normalise; uint8 RGB input only).
    """

    def __init__(self, model_identifier: str, device="cuda", *, preprocess: str = "reference",
                 image_size: int = 224, max_frames: int = 512,
                 config: Optional[ViTConfig] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("cbas_b200.DinoEncoder runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if preprocess not in ("reference", "processor"):
            raise ValueError("preprocess must be 'reference' or 'processor'")
        self.model_identifier = model_identifier
        self.preprocess = preprocess
        self.image_size = int(image_size)
        self.max_frames = int(max_frames)
        print(f"Loading DINO encoder model: {model_identifier}")
        if state_dict is not None:
            if config is None:
                raise ValueError("state_dict requires config")
            self.config, sd = config, state_dict
        elif model_identifier.startswith("synthetic:"):
            arch, _, seed = model_identifier[len("synthetic:"):].partition("@")
            if arch not in ARCHITECTURES:
                raise ValueError(f"unknown synthetic architecture '{arch}' (have {sorted(ARCHITECTURES)})")
            self.config = ViTConfig(**ARCHITECTURES[arch])
            sd = synthetic_state_dict(self.config, int(seed or 0))
        else:
            try:
                import transformers
                hf = transformers.AutoModel.from_pretrained(model_identifier)
                self.config = ViTConfig.from_hf(hf.config)
                sd = {k: v.detach() for k, v in hf.state_dict().items()}
                del hf
            except Exception as e:  # same contract as cbas.py:658-667: report, then re-raise
                print("--- MODEL LOADING FAILED ---")
                print(f"Could not load the encoder model: '{model_identifier}'. Original error: {e}")
                raise
        self._sd = {k: v.detach().cpu() for k, v in sd.items()}
        self.hidden_size = self.config.hidden_size
        self._native: Dict[Tuple[int, int, int], _NativeEncoder] = {}
        self._options: Dict[int, int] = {}
        self.eval()

    # -- construction helpers -------------------------------------------------------------------------
    @classmethod
    def from_hf_model(cls, hf_model, device="cuda", **kw) -> "DinoEncoder":
        return cls("hf-model-instance", device, config=ViTConfig.from_hf(hf_model.config),
                   state_dict={k: v.detach() for k, v in hf_model.state_dict().items()}, **kw)

    def _get_native(self, mode: int, in_hw: Tuple[int, int]) -> _NativeEncoder:
        key = (mode, in_hw[0], in_hw[1])
        nat = self._native.get(key)
        if nat is None:
            side = in_hw[0] if mode == PRE_REFERENCE else self.image_size
            if mode == PRE_REFERENCE and in_hw[0] != in_hw[1]:
                raise ValueError("reference preprocessing expects square frames (cbas.py:768-784 records square clips)")
            nat = _NativeEncoder(self.config, self._sd, self.device, mode, in_hw, side, self.max_frames)
            env = os.environ.get("CBAS_B200_LN_FUSION")  # A/B timing: 0 = standalone LayerNorm kernels, 1 = fused
            if env is not None and _lib.OPT_LN_FUSION not in self._options:
                nat.set_option(_lib.OPT_LN_FUSION, int(env))
            env = os.environ.get("CBAS_B200_SERPENTINE")  # A/B timing: 0 = every kernel walks the rows ascending
            if env is not None and _lib.OPT_SERPENTINE not in self._options:
                nat.set_option(_lib.OPT_SERPENTINE, int(env))
            for opt, val in self._options.items():
                nat.set_option(opt, val)
            self._native[key] = nat
        return nat

    def set_option(self, option: int, value: int) -> None:
        """Test knob of this encoder (cbas_b200_encoder_set_option: _lib.OPT_ATTENTION_IMPL / OPT_PRUNE_LAST_LAYER /
        OPT_RESIZE_KERNEL); applies to the handles of every geometry, present and future."""
        self._options[option] = int(value)
        for nat in self._native.values():
            nat.set_option(option, int(value))

    # -- reference-compatible call ---------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B,S,H,W] float in [0,1] (the green plane / 255, cbas.py:431,435) -> [B,S,D] fp32 on self.device."""
        B, S, H, W = x.shape
        nat = self._get_native(PRE_REFERENCE, (H, W))
        planes = x.to(self.device, dtype=torch.float32).reshape(B * S, H, W).contiguous()
        out = torch.empty(B * S, self.hidden_size, device=self.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for i in range(0, B * S, nat.max_frames):
            n = min(nat.max_frames, B * S - i)
            _lib.check(nat.lib.cbas_b200_encoder_forward_f32(
                nat.handle, planes[i:i + n].data_ptr(), n, out[i:i + n].data_ptr(), stream), "encoder_forward_f32")
        return out.reshape(B, S, self.hidden_size)

    # -- fast path used by encode_file -------------------------------------------------------------------
    def encode_u8(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: uint8 [n,H,W,3] RGB on self.device -> [n,D] fp32 CLS embeddings (asynchronous on the
        current stream)."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError("encode_u8 expects uint8 [n,H,W,3]")
        if frames.device != self.device:
            raise ValueError("encode_u8 expects frames already on the encoder's device")
        frames = frames.contiguous()
        n, H, W, _ = frames.shape
        mode = PRE_REFERENCE if self.preprocess == "reference" else PRE_PROCESSOR
        nat = self._get_native(mode, (H, W))
        if out is None:
            out = torch.empty(n, self.hidden_size, device=self.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for i in range(0, n, nat.max_frames):
            m = min(nat.max_frames, n - i)
            _lib.check(nat.lib.cbas_b200_encoder_forward_u8(
                nat.handle, frames[i:i + m].data_ptr(), m, H * W * 3, W * 3, out[i:i + m].data_ptr(), stream),
                "encoder_forward_u8")
        return out

    def encode_u8_plane(self, planes: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """planes: uint8 [n,H,W] on self.device, the green channel of each frame (what cbas.py:431 keeps) ->
        [n,D] fp32 CLS embeddings; 'reference' preprocessing only.  Same result as encode_u8 on the RGB frames."""
        if planes.dtype != torch.uint8 or planes.dim() != 3:
            raise ValueError("encode_u8_plane expects uint8 [n,H,W]")
        if self.preprocess != "reference":
            raise ValueError("single-plane input is only defined for preprocess='reference' (green / 255)")
        if planes.device != self.device:
            raise ValueError("encode_u8_plane expects planes already on the encoder's device")
        planes = planes.contiguous()
        n, H, W = planes.shape
        nat = self._get_native(PRE_REFERENCE, (H, W))
        if out is None:
            out = torch.empty(n, self.hidden_size, device=self.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for i in range(0, n, nat.max_frames):
            m = min(nat.max_frames, n - i)
            _lib.check(nat.lib.cbas_b200_encoder_forward_u8_plane(
                nat.handle, planes[i:i + m].data_ptr(), m, H * W, W, out[i:i + m].data_ptr(), stream),
                "encoder_forward_u8_plane")
        return out

    def debug_hidden(self, frames: torch.Tensor, after_layer: int) -> torch.Tensor:
        """Residual stream [n, T, D] after `after_layer` blocks (0 = embeddings); parity-test tap."""
        frames = frames.contiguous()
        n, H, W, _ = frames.shape
        mode = PRE_REFERENCE if self.preprocess == "reference" else PRE_PROCESSOR
        nat = self._get_native(mode, (H, W))
        out = torch.empty(n, nat.tokens, self.hidden_size, device=self.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(nat.lib.cbas_b200_encoder_debug_hidden(
            nat.handle, frames.data_ptr(), n, H * W * 3, W * 3, after_layer, out.data_ptr(), stream), "debug_hidden")
        return out
