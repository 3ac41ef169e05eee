"""Config presets for the models the reference trains: Llama-3-8B / Phi-3-mini-4k
+ CLIP-ViT-L/14-336, and the tiny configs the parity tests use."""
from .vlm import VisperConfig


def from_dict(c: dict, cls=VisperConfig, distill=True):
    """Build a config from the flat dict form shared with tests/bench (keys as in oracle/configs.py)."""
    vision = dict(hidden_size=c["vis_hidden"], intermediate_size=c["vis_inter"],
                  num_hidden_layers=c["vis_layers"], num_attention_heads=c["vis_heads"],
                  image_size=c["image_size"], patch_size=c["patch_size"])
    extra = {}
    if c.get("tower") == "convnext":
        vision = dict(depths=tuple(c["cnx_depths"]), dims=tuple(c["cnx_dims"]), eps=c["cnx_eps"],
                      image_size=c["image_size"])
        extra["mm_vision_tower"] = c.get("mm_vision_tower", f"CLIP-convnext_xxlarge-res{c['image_size']}")
    cfg = cls(family=c["family"], vocab_size=c["vocab"], hidden_size=c["hidden"],
              intermediate_size=c["inter"], num_hidden_layers=c["layers"], num_attention_heads=c["heads"],
              num_key_value_heads=c["kv_heads"], max_position_embeddings=c["max_pos"],
              rope_theta=c["rope_theta"], vision=vision,
              tokenizer_model_max_length=c.get("tokenizer_model_max_length", c["max_pos"]), **extra)
    if "sliding_window" in c:
        cfg.sliding_window = c["sliding_window"]
    if distill:
        li = f"d{c['depth_layers']}_s{c['seg_layers']}_g{c['gen_layers']}"
        cfg.inject_aux(mode=c.get("aux_mode", "gen-depth-seg"), layer_indices=li,
                       num_task_tokens=c.get("num_task_tokens", 8), gen_dim=c["gen_dim"],
                       seg_dim=c["seg_dim"], depth_dim=c["depth_dim"])
    return cfg


LLAMA3_8B = dict(
    family="llama", vocab=128256, hidden=4096, inter=14336, layers=32, heads=32, kv_heads=8,
    max_pos=4096, rope_theta=500000.0, vis_hidden=1024, vis_inter=4096, vis_layers=24, vis_heads=16,
    image_size=336, patch_size=14, gen_dim=1024, seg_dim=1536, depth_dim=1024, depth_layers="18-20",
    seg_layers="10-18", gen_layers="12-20", aux_mode="gen-depth-seg", num_task_tokens=8,
    tokenizer_model_max_length=4096)

PHI3_MINI = dict(
    family="phi3", vocab=32064, hidden=3072, inter=8192, layers=32, heads=32, kv_heads=32,
    max_pos=4096, rope_theta=10000.0, vis_hidden=1024, vis_inter=4096, vis_layers=24, vis_heads=16,
    image_size=336, patch_size=14, gen_dim=1024, seg_dim=1536, depth_dim=1024, depth_layers="18-20",
    seg_layers="10-18", gen_layers="12-20", aux_mode="gen-depth-seg", num_task_tokens=8,
    tokenizer_model_max_length=4096)
