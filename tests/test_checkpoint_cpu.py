"""Host-side checkpoint ingest (lightdiffusion_next_b200/checkpoint.py): container read, SD1.5 prefix split + layout
validation, LoRA merge arithmetic (reference: Loader.py:11-111, SD15.py:32-69, LoRas.py:15-121, ModelPatcher.py:621-650)."""
import os

import pytest
import torch

from lightdiffusion_next_b200 import checkpoint as C
from lightdiffusion_next_b200 import synth


def test_safetensors_roundtrip_and_prefix_split(tmp_path):
    from safetensors.torch import save_file
    g = torch.Generator().manual_seed(0)
    sd = {
        "model.diffusion_model.time_embed.0.weight": torch.randn(8, 4, generator=g).half(),
        "model.diffusion_model.out.2.bias": torch.randn(4, generator=g).half(),
        "first_stage_model.decoder.conv_in.weight": torch.randn(8, 4, 3, 3, generator=g).half(),
        "first_stage_model.post_quant_conv.bias": torch.randn(4, generator=g).half(),
        "first_stage_model.encoder.conv_in.weight": torch.randn(8, 3, 3, 3, generator=g).half(),  # encoder side (img2img)
        # old CLIP layout (no `text_model.`), renamed like SD15.process_clip_state_dict
        "cond_stage_model.transformer.embeddings.position_embedding.weight": torch.randn(77, 8, generator=g).half(),
        "cond_stage_model.transformer.text_model.final_layer_norm.bias": torch.randn(8, generator=g).half(),
        "cond_stage_model.transformer.text_model.embeddings.position_ids": torch.arange(77)[None],
        "model_ema.decay": torch.tensor(0.999),
    }
    path = str(tmp_path / "tiny.safetensors")
    save_file(sd, path)
    loaded = C.load_state_dict_file(path)
    assert set(loaded) == set(sd)
    parts = C.split_sd15_checkpoint(loaded, strict=False)
    assert set(parts["unet"]) == {"time_embed.0.weight", "out.2.bias"}
    assert set(parts["vae"]) == {"decoder.conv_in.weight", "post_quant_conv.bias", "encoder.conv_in.weight"}
    assert set(parts["clip"]) == {"embeddings.position_embedding.weight", "final_layer_norm.bias"}
    assert torch.equal(parts["unet"]["time_embed.0.weight"], sd["model.diffusion_model.time_embed.0.weight"])
    # pickled container with a top-level "state_dict"
    p2 = str(tmp_path / "tiny.ckpt")
    torch.save({"state_dict": sd}, p2)
    assert set(C.load_state_dict_file(p2)) == set(sd)


def _meta_checkpoint():
    sd = {}
    for k, s in synth.unet_shapes().items():
        sd[C.UNET_PREFIX + k] = torch.empty(s, dtype=torch.float16, device="meta")
    for k, s in synth.vae_decoder_shapes().items():
        sd[C.VAE_PREFIX + k] = torch.empty(s, dtype=torch.float16, device="meta")
    for k, s in synth.clip_shapes().items():
        sd[C.CLIP_PREFIXES[0] + k] = torch.empty(s, dtype=torch.float16, device="meta")
    return sd


def test_strict_layout_validation():
    sd = _meta_checkpoint()
    parts = C.split_sd15_checkpoint(sd)
    assert len(parts["unet"]) == len(synth.unet_shapes()) == 686
    assert len(parts["vae"]) == len(synth.vae_decoder_shapes())
    assert len(parts["clip"]) == len(synth.clip_shapes())
    # a 1x1 conv stored as a linear weight is accepted and reshaped
    k = C.UNET_PREFIX + "input_blocks.1.1.proj_in.weight"
    sd2 = dict(sd)
    sd2[k] = torch.empty(320, 320, dtype=torch.float16, device="meta")
    assert tuple(C.split_sd15_checkpoint(sd2)["unet"]["input_blocks.1.1.proj_in.weight"].shape) == (320, 320, 1, 1)
    # missing / mis-shaped tensors are reported by name
    sd3 = dict(sd)
    del sd3[C.UNET_PREFIX + "middle_block.1.transformer_blocks.0.attn2.to_k.weight"]
    sd3[C.VAE_PREFIX + "decoder.conv_out.weight"] = torch.empty(3, 64, 3, 3, dtype=torch.float16, device="meta")
    with pytest.raises(ValueError) as ei:
        C.split_sd15_checkpoint(sd3)
    msg = str(ei.value)
    assert "middle_block.1.transformer_blocks.0.attn2.to_k.weight" in msg and "decoder.conv_out.weight" in msg
    # with the encoder side present it is validated too (img2img needs it); decode-only VAE files stay valid
    sd4 = dict(sd)
    for k, shp in synth.vae_encoder_shapes().items():
        sd4[C.VAE_PREFIX + k] = torch.empty(shp, dtype=torch.float16, device="meta")
    assert len(C.split_sd15_checkpoint(sd4)["vae"]) == len(synth.vae_decoder_shapes()) + len(synth.vae_encoder_shapes())
    del sd4[C.VAE_PREFIX + "quant_conv.weight"]
    with pytest.raises(ValueError, match="quant_conv.weight"):
        C.split_sd15_checkpoint(sd4)
    # a UNet-only file loads (other parts empty)
    only = {k: v for k, v in sd.items() if k.startswith(C.UNET_PREFIX)}
    p = C.split_sd15_checkpoint(only)
    assert p["vae"] == {} and p["clip"] == {} and len(p["unet"]) == 686


def test_merge_lora_matches_reference_formula():
    g = torch.Generator().manual_seed(1)
    parts = {
        "unet": {
            "input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight": torch.randn(32, 32, generator=g).half(),
            "input_blocks.1.0.in_layers.2.weight": torch.randn(16, 8, 3, 3, generator=g).half(),
            "input_blocks.1.0.in_layers.2.bias": torch.randn(16, generator=g).half(),
        },
        "clip": {"encoder.layers.3.self_attn.k_proj.weight": torch.randn(24, 24, generator=g).half(),
                 "encoder.layers.3.mlp.fc1.weight": torch.randn(48, 24, generator=g).half()},
        "vae": {},
    }
    ref = {p: {k: v.clone() for k, v in d.items()} for p, d in parts.items()}
    r = 4
    lora = {
        "lora_unet_input_blocks_1_1_transformer_blocks_0_attn1_to_q.lora_up.weight": torch.randn(32, r, generator=g),
        "lora_unet_input_blocks_1_1_transformer_blocks_0_attn1_to_q.lora_down.weight": torch.randn(r, 32, generator=g),
        "lora_unet_input_blocks_1_1_transformer_blocks_0_attn1_to_q.alpha": torch.tensor(2.0),
        "lora_unet_input_blocks_1_0_in_layers_2.lora_up.weight": torch.randn(16, r, 1, 1, generator=g),
        "lora_unet_input_blocks_1_0_in_layers_2.lora_down.weight": torch.randn(r, 8, 3, 3, generator=g),
        "lora_te_text_model_encoder_layers_3_self_attn_k_proj.lora_up.weight": torch.randn(24, r, generator=g),
        "lora_te_text_model_encoder_layers_3_self_attn_k_proj.lora_down.weight": torch.randn(r, 24, generator=g),
        "lora_te_text_model_encoder_layers_3_self_attn_k_proj.alpha": torch.tensor(4.0),
        "lora_unet_some_module_not_in_the_model.lora_up.weight": torch.randn(4, r, generator=g),
        "lora_unet_some_module_not_in_the_model.lora_down.weight": torch.randn(r, 4, generator=g),
    }
    n = C.merge_lora(parts, lora, strength_model=0.8, strength_clip=0.5)
    assert n == 3

    def expect(w, up, down, strength, alpha):
        a = strength * (alpha / down.shape[0] if alpha is not None else 1.0)
        return (w.float() + (a * torch.mm(up.float().flatten(1), down.float().flatten(1))).reshape(w.shape)).to(w.dtype)

    k = "input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight"
    m = "lora_unet_input_blocks_1_1_transformer_blocks_0_attn1_to_q"
    assert torch.equal(parts["unet"][k], expect(ref["unet"][k], lora[m + ".lora_up.weight"], lora[m + ".lora_down.weight"], 0.8, 2.0))
    k = "input_blocks.1.0.in_layers.2.weight"
    m = "lora_unet_input_blocks_1_0_in_layers_2"
    assert torch.equal(parts["unet"][k], expect(ref["unet"][k], lora[m + ".lora_up.weight"], lora[m + ".lora_down.weight"], 0.8, None))
    k = "encoder.layers.3.self_attn.k_proj.weight"
    m = "lora_te_text_model_encoder_layers_3_self_attn_k_proj"
    assert torch.equal(parts["clip"][k], expect(ref["clip"][k], lora[m + ".lora_up.weight"], lora[m + ".lora_down.weight"], 0.5, 4.0))
    # untouched tensors stay bit-identical
    assert torch.equal(parts["unet"]["input_blocks.1.0.in_layers.2.bias"], ref["unet"]["input_blocks.1.0.in_layers.2.bias"])
    assert torch.equal(parts["clip"]["encoder.layers.3.mlp.fc1.weight"], ref["clip"]["encoder.layers.3.mlp.fc1.weight"])
    # zero strength is a no-op
    again = {p: {k: v.clone() for k, v in d.items()} for p, d in ref.items()}
    assert C.merge_lora(again, lora, 0.0, 0.0) == 0


def test_gguf_q8_0_ingest(tmp_path):
    """GGUF reader: Q8_0 blocks are dequantised once (w = d * q, the reference's dequantize_blocks_Q8_0 arithmetic), F32 /
    F16 pass through, the model prefix is stripped like gguf_sd_loader does, unknown quantisations are rejected."""
    import gguf
    import numpy as np
    rng = np.random.default_rng(0)
    w_q = rng.standard_normal((48, 64)).astype(np.float32)          # rows of 64 = 2 Q8_0 blocks each
    w_h = rng.standard_normal((8, 16)).astype(np.float16)
    b_f = rng.standard_normal((48,)).astype(np.float32)
    path = str(tmp_path / "tiny.gguf")
    wr = gguf.GGUFWriter(path, "flux")
    q = gguf.quants.quantize(w_q, gguf.GGMLQuantizationType.Q8_0)
    wr.add_tensor("model.diffusion_model.double_blocks.0.img_attn.proj.weight", q, raw_dtype=gguf.GGMLQuantizationType.Q8_0)
    wr.add_tensor("model.diffusion_model.img_in.weight", w_h)
    wr.add_tensor("model.diffusion_model.img_in.bias", b_f)
    wr.add_tensor("some.other.tensor", b_f)
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    sd = C.load_gguf(path)
    assert set(sd) == {"double_blocks.0.img_attn.proj.weight", "img_in.weight", "img_in.bias"}
    ref = torch.from_numpy(gguf.quants.dequantize(q, gguf.GGMLQuantizationType.Q8_0))
    got = sd["double_blocks.0.img_attn.proj.weight"]
    assert got.dtype == torch.bfloat16 and got.shape == (48, 64)
    assert torch.equal(got, ref.to(torch.bfloat16))                 # same d * q product, one rounding to bf16
    assert (got.float() - torch.from_numpy(w_q)).abs().max() < 0.05  # and it is the quantised original
    assert torch.equal(sd["img_in.weight"], torch.from_numpy(w_h)) and sd["img_in.weight"].dtype == torch.float16
    assert torch.equal(sd["img_in.bias"], torch.from_numpy(b_f))
    # through the generic entry point (no prefix handling) every tensor is returned
    assert len(C.load_state_dict_file(path)) == 4
    # an unsupported quantisation is refused loudly
    p2 = str(tmp_path / "q4.gguf")
    wr = gguf.GGUFWriter(p2, "flux")
    wr.add_tensor("w", gguf.quants.quantize(w_q, gguf.GGMLQuantizationType.Q4_0), raw_dtype=gguf.GGMLQuantizationType.Q4_0)
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    with pytest.raises(ValueError, match="unsupported type"):
        C.load_gguf(p2)


def test_t5_gguf_key_map(tmp_path):
    """llama.cpp T5 encoder names -> the reference's T5 state-dict keys (clip_sd_map, Quantizer.py:815-858): the mapped dict
    has exactly the layout Engine.load_t5 expects; Q8_0 matrices are dequantised once, F32 norms pass through."""
    import gguf
    import numpy as np
    from lightdiffusion_next_b200 import t5 as T5H
    shapes = T5H.t5_shapes(d_model=64, d_ff=96, num_heads=1, num_layers=2, vocab_size=40)
    inv = {"shared": "token_embd", "encoder.final_layer_norm": "enc.output_norm"}
    sub = {"layer.0.SelfAttention.q": "attn_q", "layer.0.SelfAttention.k": "attn_k", "layer.0.SelfAttention.v": "attn_v",
           "layer.0.SelfAttention.o": "attn_o", "layer.0.layer_norm": "attn_norm",
           "layer.0.SelfAttention.relative_attention_bias": "attn_rel_b", "layer.1.DenseReluDense.wi_1": "ffn_up",
           "layer.1.DenseReluDense.wo": "ffn_down", "layer.1.DenseReluDense.wi_0": "ffn_gate", "layer.1.layer_norm": "ffn_norm"}
    rng = np.random.default_rng(1)
    path = str(tmp_path / "t5.gguf")
    wr = gguf.GGUFWriter(path, "t5encoder")
    want = {}
    for key, shape in shapes.items():
        stem = key[:-len(".weight")]
        if stem in inv:
            name = inv[stem]
        else:
            blk, rest = stem[len("encoder.block."):].split(".", 1)
            name = f"enc.blk.{blk}.{sub[rest]}"
        w = rng.standard_normal(shape).astype(np.float32)
        if len(shape) == 2 and shape[1] % 32 == 0:
            q = gguf.quants.quantize(w, gguf.GGMLQuantizationType.Q8_0)
            wr.add_tensor(name + ".weight", q, raw_dtype=gguf.GGMLQuantizationType.Q8_0)
            want[key] = torch.from_numpy(gguf.quants.dequantize(q, gguf.GGMLQuantizationType.Q8_0)).to(torch.bfloat16)
        else:
            wr.add_tensor(name + ".weight", w)
            want[key] = torch.from_numpy(w)
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    sd = C.load_t5_gguf(path)
    assert {k: tuple(v.shape) for k, v in sd.items()} == shapes
    for k in shapes:
        assert torch.equal(sd[k], want[k]), k
    # a GGUF that is not a T5 encoder is refused (the reference asserts on enc.blk.23.ffn_up.weight)
    p2 = str(tmp_path / "other.gguf")
    wr = gguf.GGUFWriter(p2, "flux")
    wr.add_tensor("img_in.bias", rng.standard_normal((8,)).astype(np.float32))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    with pytest.raises(ValueError, match="not a T5 encoder"):
        C.load_t5_gguf(p2)


def test_t5_gguf_key_map_matches_reference_map():
    """Every llama.cpp T5 name maps exactly as the reference's own clip_sd_map maps it (fixture from make_golden_t5.py)."""
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "t5_small.pt"))
    from lightdiffusion_next_b200 import t5 as T5H
    xxl = T5H.t5_shapes()
    for name, want in gold["gguf_name_map"].items():
        got = name
        for a, b in C.T5_GGUF_KEY_MAP:
            got = got.replace(a, b)
        assert got == want, (name, got, want)
        assert want in xxl or "relative_attention_bias" in want  # only block 0 owns a bias table



def test_load_flux_files_routes_every_part(tmp_path):
    """checkpoint.load_flux_files: DiT from GGUF (prefix stripped, Q8_0 dequantised), autoencoder / CLIP-L from safetensors
    (HF `text_model.` prefix stripped, position_ids dropped), T5 from GGUF (llama.cpp names mapped) -> the engine's loaders."""
    import gguf
    import numpy as np
    from safetensors.torch import save_file
    rng = np.random.default_rng(2)
    unet = str(tmp_path / "flux.gguf")
    wr = gguf.GGUFWriter(unet, "flux")
    w = rng.standard_normal((8, 64)).astype(np.float32)
    wr.add_tensor("model.diffusion_model.img_in.weight", gguf.quants.quantize(w, gguf.GGMLQuantizationType.Q8_0),
                  raw_dtype=gguf.GGMLQuantizationType.Q8_0)
    wr.add_tensor("model.diffusion_model.img_in.bias", rng.standard_normal((8,)).astype(np.float32))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    ae = str(tmp_path / "ae.safetensors")
    save_file({"decoder.conv_in.weight": torch.zeros(4, 16, 3, 3), "encoder.conv_in.bias": torch.zeros(4),
               "loss.logvar": torch.zeros(1)}, ae)
    clip = str(tmp_path / "clip_l.safetensors")
    save_file({"text_model.embeddings.token_embedding.weight": torch.zeros(10, 8), "text_model.embeddings.position_ids": torch.zeros(1, 77),
               "text_model.encoder.layers.0.mlp.fc1.weight": torch.zeros(4, 8), "text_model.final_layer_norm.bias": torch.zeros(8),
               "text_projection.weight": torch.zeros(8, 8)}, clip)
    t5 = str(tmp_path / "t5.gguf")
    wr = gguf.GGUFWriter(t5, "t5encoder")
    wr.add_tensor("token_embd.weight", rng.standard_normal((6, 32)).astype(np.float32))
    wr.add_tensor("enc.blk.0.ffn_up.weight", rng.standard_normal((8, 32)).astype(np.float32))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()

    class Rec:
        def __init__(self):
            self.got = {}

        def load_flux(self, sd): self.got["flux"] = sd
        def load_vae(self, sd): self.got["vae"] = sd
        def load_clip(self, sd): self.got["clip"] = sd
        def load_t5(self, sd): self.got["t5"] = sd

    e = Rec()
    n = C.load_flux_files(e, unet, ae, clip, t5)
    assert n == {"flux": 2, "vae": 2, "clip": 3, "t5": 2}
    assert set(e.got["flux"]) == {"img_in.weight", "img_in.bias"} and e.got["flux"]["img_in.weight"].dtype == torch.bfloat16
    assert set(e.got["vae"]) == {"decoder.conv_in.weight", "encoder.conv_in.bias"}
    assert set(e.got["clip"]) == {"embeddings.token_embedding.weight", "encoder.layers.0.mlp.fc1.weight", "final_layer_norm.bias"}
    assert set(e.got["t5"]) == {"shared.weight", "encoder.block.0.layer.1.DenseReluDense.wi_1.weight"}
    with pytest.raises(ValueError, match="not a CLIP-L"):
        C.load_flux_files(Rec(), unet, clip_l_path=ae)
    with pytest.raises(ValueError, match="not a Flux autoencoder"):
        C.load_flux_files(Rec(), unet, ae_path=clip)


def test_gguf_orig_shape_metadata_and_architecture_check(tmp_path):
    """gguf_sd_loader behaviours (Quantizer.py:427-447, 600-614): `comfy.gguf.orig_shape.<name>` restores a tensor's original
    shape (converters flatten some tensors to quantise them); an unknown `general.architecture` is refused."""
    import gguf
    import numpy as np
    rng = np.random.default_rng(3)
    w = rng.standard_normal((4, 64)).astype(np.float32)            # stored flat-ish, really [4, 16, 2, 2]
    path = str(tmp_path / "shaped.gguf")
    wr = gguf.GGUFWriter(path, "sd1")
    wr.add_tensor("model.diffusion_model.input_blocks.0.0.weight", gguf.quants.quantize(w, gguf.GGMLQuantizationType.Q8_0),
                  raw_dtype=gguf.GGMLQuantizationType.Q8_0)
    wr.add_array("comfy.gguf.orig_shape.model.diffusion_model.input_blocks.0.0.weight", [4, 16, 2, 2])
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    sd = C.load_gguf(path)
    got = sd["input_blocks.0.0.weight"]
    assert got.shape == (4, 16, 2, 2)
    ref = torch.from_numpy(gguf.quants.dequantize(gguf.quants.quantize(w, gguf.GGMLQuantizationType.Q8_0), gguf.GGMLQuantizationType.Q8_0))
    assert torch.equal(got, ref.reshape(4, 16, 2, 2).to(torch.bfloat16))
    p2 = str(tmp_path / "llama.gguf")
    wr = gguf.GGUFWriter(p2, "llama")
    wr.add_tensor("w", w)
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    with pytest.raises(ValueError, match="unexpected GGUF architecture"):
        C.load_gguf(p2)


def test_lora_unet_key_map_equals_the_reference_map():
    """Every LoRA name the reference resolves to an existing SD1.5 UNet weight (model_lora_keys_unet incl. the diffusers names
    kohya LoRAs use; tests/golden/make_golden_lora_keys.py) resolves to the same weight here, and nothing else is claimed;
    the LDM -> diffusers naming is the exact inverse of the reference's unet_to_diffusers table."""
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lora_keys.pt"))
    unet_keys = {k[len("diffusion_model."):] for k in gold["unet_keys"]}
    assert unet_keys == set(synth.unet_shapes())
    want = {name: tgt[len("diffusion_model."):] for name, tgt in gold["lora_to_unet"].items() if tgt in set(gold["unet_keys"])}
    got = {name: key for name, (part, key) in C.lora_key_map({"unet": synth.unet_shapes(), "clip": {}}).items() if part == "unet"}
    assert got == want
    assert "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q" in got
    assert got["lora_unet_up_blocks_3_resnets_2_conv_shortcut"] == "output_blocks.11.0.skip_connection.weight"
    inv = {ldm: d for d, ldm in gold["diffusers_map"].items() if ldm in unet_keys}
    for k in unet_keys:
        assert C.diffusers_unet_name(k) == inv.get(k), k


def test_merge_lora_through_diffusers_names():
    """A kohya-style SD1.5 LoRA names UNet modules by their diffusers path: the merge lands on the LDM weight the reference's
    key map points at, with the same arithmetic as for LDM-named modules."""
    g = torch.Generator().manual_seed(5)
    k = "output_blocks.4.1.transformer_blocks.0.attn2.to_k.weight"      # = up_blocks.1.attentions.1 in diffusers
    w0 = torch.randn(16, 24, generator=g).half()
    parts = {"unet": {k: w0.clone()}, "clip": {}, "vae": {}}
    m = "lora_unet_up_blocks_1_attentions_1_transformer_blocks_0_attn2_to_k"
    up, down = torch.randn(16, 2, generator=g), torch.randn(2, 24, generator=g)
    assert C.merge_lora(parts, {m + ".lora_up.weight": up, m + ".lora_down.weight": down, m + ".alpha": torch.tensor(1.0)}, 0.7) == 1
    assert torch.equal(parts["unet"][k], (w0.float() + 0.7 * (1.0 / 2) * torch.mm(up, down)).to(w0.dtype))


def test_lora_application_is_bit_identical_to_the_reference_patcher():
    """merge_lora (ingest path) and backend.unet_state_dict_from_model (seam path, patches queued on a ModelPatcher) against
    weights patched by the reference's own LoRas.load_lora + ModelPatcher.add_patches / patch_model
    (tests/golden/make_golden_lora_apply.py): same module-name resolution, same fp32 sum rounded once to fp16 -- bit for bit."""
    import types
    from lightdiffusion_next_b200 import backend
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lora_apply.pt"))
    keys = list(gold["patched"])
    base = {k: synth.synth_tensor(k, synth.unet_shapes()[k]) for k in keys}
    parts = {"unet": {k: v.clone() for k, v in base.items()}, "clip": {}, "vae": {}}
    assert C.merge_lora(parts, gold["lora"], strength_model=gold["strength"]) == 3
    for k in keys:
        assert not torch.equal(parts["unet"][k], base[k])
        assert torch.equal(parts["unet"][k], gold["patched"][k]), k

    # seam path: a ModelPatcher-like object with queued patches in the reference's own tuple format
    def calculate_weight(patches, weight, key):                     # stand-in with the reference's signature
        for strength, (kind, (up, down, alpha, mid, dora)), _ in patches:
            a = strength * (alpha / down.shape[0] if alpha is not None else 1.0)
            weight += (a * torch.mm(up.float().flatten(1), down.float().flatten(1))).reshape(weight.shape)
        return weight

    km = {m: k for m, (p, k) in C.lora_key_map({"unet": synth.unet_shapes(), "clip": {}}).items()}
    patches = {}
    for name in {n.rsplit(".lora_up.weight", 1)[0] for n in gold["lora"] if n.endswith(".lora_up.weight")}:
        al = gold["lora"].get(name + ".alpha")
        patches["diffusion_model." + km[name]] = [(gold["strength"], ("lora", (gold["lora"][name + ".lora_up.weight"],
                                                   gold["lora"][name + ".lora_down.weight"], None if al is None else al.item(), None, None)), 1.0)]
    dm = types.SimpleNamespace(state_dict=lambda: {k: v.clone() for k, v in base.items()})
    patcher = types.SimpleNamespace(patches=patches, calculate_weight=calculate_weight)
    sd = backend.unet_state_dict_from_model(types.SimpleNamespace(diffusion_model=dm), patcher)
    for k in keys:
        assert torch.equal(sd[k], gold["patched"][k]), k
    assert all(torch.equal(a, b) for a, b in zip(dm.state_dict().values(), base.values()))   # the source module is untouched


def test_clip_patches_are_folded_in_at_the_seam():
    """backend.clip_state_dict_from_clip: the CLIP half of a LoRA queued on the reference's `clip.patcher` lands on the
    text-model weights handed to Engine.load_clip (same arithmetic as the UNet path); position_ids is dropped."""
    import types
    from lightdiffusion_next_b200 import backend
    g = torch.Generator().manual_seed(2)
    w = torch.randn(24, 24, generator=g).half()
    up, down = torch.randn(24, 2, generator=g), torch.randn(2, 24, generator=g)

    def calculate_weight(patches, weight, key):
        for strength, (kind, (u, d, alpha, mid, dora)), _ in patches:
            weight += strength * (alpha / d.shape[0]) * torch.mm(u.float(), d.float())
        return weight

    tm = types.SimpleNamespace(state_dict=lambda: {"encoder.layers.0.self_attn.q_proj.weight": w.clone(),
                                                   "embeddings.position_ids": torch.arange(77)[None]})
    clip = types.SimpleNamespace(
        cond_stage_model=types.SimpleNamespace(clip_l=types.SimpleNamespace(transformer=types.SimpleNamespace(text_model=tm))),
        patcher=types.SimpleNamespace(calculate_weight=calculate_weight, patches={
            "clip_l.transformer.text_model.encoder.layers.0.self_attn.q_proj.weight": [(0.7, ("lora", (up, down, 4.0, None, None)), 1.0)],
            "clip_l.logit_scale": [(1.0, ("lora", (up, down, 1.0, None, None)), 1.0)]}))
    sd = backend.clip_state_dict_from_clip(clip)
    assert set(sd) == {"encoder.layers.0.self_attn.q_proj.weight"}
    assert torch.equal(sd["encoder.layers.0.self_attn.q_proj.weight"], (w.float() + 0.7 * (4.0 / 2) * torch.mm(up, down)).half())
