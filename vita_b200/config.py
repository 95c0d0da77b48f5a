"""Geometry of the omni forward path.

Full-size values come from the reference's shipped config
(web_demo/vllm_tools/model_weight_file/config.json:16-109): Mixtral-8x7B text model with a 51760-entry vocabulary,
InternViT-300M-448px, Whale audio encoder + CNNSubsampling adapter.  `tiny()` keeps every kernel-relevant constant
(head dims 128 / 64 / 64, 8 experts top-2, patch 14, 80-bin fbank) and shrinks widths/depths so the CPU oracle and
the golden fixtures stay small.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict

IMAGE_TOKEN_INDEX = -200   # vita/constants.py:5
AUDIO_TOKEN_INDEX = -500   # vita/constants.py:6
IGNORE_INDEX = -100        # vita/constants.py:4


@dataclass
class LLMConfig:
    vocab_size: int = 51760
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 128
    num_local_experts: int = 8
    num_experts_per_tok: int = 2
    rms_norm_eps: float = 1e-5
    rope_theta: float = 1e6
    max_position_embeddings: int = 32768
    tokenizer_model_max_length: int = 4600   # config.json:116, applied at vita_arch.py:326-329

    @property
    def qkv_rows(self) -> int:
        return (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim


@dataclass
class VisionConfig:
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    image_size: int = 448
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-6
    scale_pix_shuffle: float = 0.5     # internvit_encoder.py:16

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def patch_k(self) -> int:
        return self.num_channels * self.patch_size * self.patch_size

    @property
    def patch_k_pad(self) -> int:          # im2col row length, 16-byte aligned for TMA
        return (self.patch_k + 7) // 8 * 8

    @property
    def out_tokens(self) -> int:           # tokens per tile after pixel shuffle
        return self.num_patches // 4

    @property
    def out_dim(self) -> int:              # internvit_encoder.py:100-102
        return self.hidden_size * 4


@dataclass
class AudioConfig:
    input_dim: int = 80
    hidden_size: int = 1024
    num_attention_heads: int = 16
    linear_units: int = 4096
    num_blocks: int = 24
    adapter_kernel: int = 5
    adapter_ln_eps: float = 1e-3          # adapter.py:98
    layer_norm_eps: float = 1e-5          # torch.nn.LayerNorm default
    max_len: int = 5000                   # attention.py:88

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def freq_bins(self) -> int:           # ((idim - 1) // 2 - 1) // 2, subsampling.py:34
        return ((self.input_dim - 1) // 2 - 1) // 2

    @staticmethod
    def frames_after_subsampling(t: int) -> int:   # two Conv2d(k=3, s=2)
        return ((t - 1) // 2 - 1) // 2

    @staticmethod
    def tokens_after_adapter(t2: int) -> int:      # right-pad k-1, Conv1d(k, stride 2); init_model.py:57-58
        return (t2 - 1) // 2 + 1


@dataclass
class VitaConfig:
    llm: LLMConfig = field(default_factory=LLMConfig)
    vision: VisionConfig = field(default_factory=VisionConfig)
    audio: AudioConfig = field(default_factory=AudioConfig)
    mm_projector_type: str = "mlp2x_gelu"   # multimodal_projector/builder.py:160-168

    def to_dict(self):
        return asdict(self)

    @staticmethod
    def full(num_hidden_layers: int = 32) -> "VitaConfig":
        return VitaConfig(llm=LLMConfig(num_hidden_layers=num_hidden_layers))

    @staticmethod
    def tiny() -> "VitaConfig":
        return VitaConfig(
            llm=LLMConfig(vocab_size=2048, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                          num_attention_heads=4, num_key_value_heads=1, max_position_embeddings=2048,
                          tokenizer_model_max_length=1024),
            vision=VisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                                image_size=112),
            audio=AudioConfig(hidden_size=128, num_attention_heads=2, linear_units=256, num_blocks=2, max_len=1000),
        )
