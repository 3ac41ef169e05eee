"""Model configurations shared by the oracle, the golden-vector generator, tests and bench.py.
Full-size numbers follow SURVEY.md §8 (Llama-3-8B, Phi-3-mini-4k, CLIP-ViT-L/14-336) and the
distill head defaults of /root/reference/ola_vlm/train/ola_vlm_train.py:84-107 with
layer_indices d18-20_s10-18_g12-20 (scripts/train/pretrain.sh:19-23)."""

TINY_LLAMA = dict(
    family="llama", vocab=512, hidden=128, inter=256, layers=4, heads=4, kv_heads=2, max_pos=1024,
    rope_theta=500000.0, vis_hidden=64, vis_inter=128, vis_layers=3, vis_heads=2, image_size=336,
    patch_size=14, gen_dim=64, seg_dim=96, depth_dim=64, depth_layers="3-4", seg_layers="1-3",
    gen_layers="2-4", aux_mode="gen-depth-seg", num_task_tokens=8, num_sys_tokens=26,
    tokenizer_model_max_length=1024)

# BASELINE config 4 at test size: the CLIP-ConvNeXt tower (768 px → 24x24 = 576 image tokens, like the ViT)
# in front of the tiny Llama; widths are multiples of 64 (one 128-byte line per pixel in the depthwise kernel)
TINY_LLAMA_CONVNEXT = dict(TINY_LLAMA, tower="convnext", cnx_depths=(1, 1, 2, 1), cnx_dims=(64, 64, 128, 64),
                           cnx_eps=1e-5, image_size=768)

TINY_PHI3 = dict(TINY_LLAMA, family="phi3", kv_heads=4, rope_theta=10000.0, num_sys_tokens=13)
# Phi-3's sliding-window attention (config 5: T=4096 > sliding_window 2047) at test size: the
# embedded sequence is ~650 tokens, so a 200-token window cuts into image, task and text tokens.
TINY_PHI3_SW = dict(TINY_PHI3, sliding_window=200)

# real layer WIDTHS (Llama-3-8B / Phi-3-mini / CLIP-ViT-L dims, all three head widths) at reduced depth
# and vocabulary: exercises the production kernel shapes (hd 128 GQA 32/8, hd 96, K=14336, dim-4096
# depth head) while the fp32 CPU oracle still finishes in seconds.
WIDE_LLAMA = dict(
    family="llama", vocab=8192, hidden=4096, inter=14336, layers=2, heads=32, kv_heads=8, max_pos=1024,
    rope_theta=500000.0, vis_hidden=1024, vis_inter=4096, vis_layers=3, vis_heads=16, image_size=336,
    patch_size=14, gen_dim=1024, seg_dim=1536, depth_dim=1024, depth_layers="1-2", seg_layers="1-2",
    gen_layers="1-2", aux_mode="gen-depth-seg", num_task_tokens=8, num_sys_tokens=26,
    tokenizer_model_max_length=1024)

WIDE_PHI3 = dict(WIDE_LLAMA, family="phi3", hidden=3072, inter=8192, heads=32, kv_heads=32,
                 rope_theta=10000.0, num_sys_tokens=13)

LLAMA3_8B = dict(
    family="llama", vocab=128256, hidden=4096, inter=14336, layers=32, heads=32, kv_heads=8,
    max_pos=4096, rope_theta=500000.0, vis_hidden=1024, vis_inter=4096, vis_layers=24, vis_heads=16,
    image_size=336, patch_size=14, gen_dim=1024, seg_dim=1536, depth_dim=1024, depth_layers="18-20",
    seg_layers="10-18", gen_layers="12-20", aux_mode="gen-depth-seg", num_task_tokens=8,
    num_sys_tokens=38, tokenizer_model_max_length=4096)

PHI3_MINI = dict(
    family="phi3", vocab=32064, hidden=3072, inter=8192, layers=32, heads=32, kv_heads=32,
    max_pos=4096, rope_theta=10000.0, vis_hidden=1024, vis_inter=4096, vis_layers=24, vis_heads=16,
    image_size=336, patch_size=14, gen_dim=1024, seg_dim=1536, depth_dim=1024, depth_layers="18-20",
    seg_layers="10-18", gen_layers="12-20", aux_mode="gen-depth-seg", num_task_tokens=8,
    num_sys_tokens=13, tokenizer_model_max_length=4096)


def synthetic_batch_mixed(cfg, n_text, seed=1234):
    """Modality edge cases of the splice (ola_arch.py:345-391) for the NTP-only classes: row 0 has
    one image, row 1 is TEXT-ONLY (still consumes an image slot, :348-355) and is right-padded, row 2
    has TWO images.  `images` therefore carries 4 images for 3 rows."""
    import torch

    g = torch.Generator().manual_seed(seed)
    S, V = cfg["num_sys_tokens"], cfg["vocab"]
    B = 3
    ids = torch.randint(0, V - 1, (B, n_text), generator=g)
    ids[0, S] = -200
    ids[2, S] = -200
    ids[2, S + 5] = -200
    labels = ids.clone()
    labels[:, :S + 8] = -100
    am = torch.ones(B, n_text, dtype=torch.bool)
    keep = (2 * n_text) // 3
    ids[1, keep:] = V - 1
    labels[1, keep:] = -100
    am[1, keep:] = False
    images = torch.randn(4, 3, cfg["image_size"], cfg["image_size"], generator=g)
    return dict(input_ids=ids, labels=labels, attention_mask=am, images=images)


def synthetic_batch(cfg, B, n_text, seed=1234, distill=True, pad_rows=0, dtype=None):
    """SURVEY.md §8(d) synthetic batch: one IMAGE_TOKEN_INDEX at position S, labels = ids with the
    first S+8 positions ignored, N(0,1) images/targets, ones masks (int64 as the collator makes them).
    pad_rows > 0 right-pads that many trailing rows to 3/4 length (varlen variant)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    S, V = cfg["num_sys_tokens"], cfg["vocab"]
    ids = torch.randint(0, V - 1, (B, n_text), generator=g)
    ids[:, S] = -200
    labels = ids.clone()
    labels[:, :S + 8] = -100
    am = torch.ones(B, n_text, dtype=torch.bool)
    pad_id = V - 1
    for r in range(B - pad_rows, B):
        keep = (3 * n_text) // 4
        ids[r, keep:] = pad_id
        labels[r, keep:] = -100
        am[r, keep:] = False
    images = torch.randn(B, 3, cfg["image_size"], cfg["image_size"], generator=g)
    batch = dict(input_ids=ids, labels=labels, attention_mask=am, images=images)
    if distill:
        batch["targets"] = dict(
            depth=torch.randn(B, 576, cfg["depth_dim"], generator=g),
            seg=torch.randn(B, cfg["seg_dim"], 24, 24, generator=g),
            gen=torch.randn(B, 1, cfg["gen_dim"], generator=g))
        batch["masks"] = dict(depth=torch.ones(B, dtype=torch.long), seg=torch.ones(B, dtype=torch.long),
                              gen=torch.ones(B, dtype=torch.long))
    if dtype is not None:
        batch["images"] = batch["images"].to(dtype)
        if distill:
            batch["targets"] = {k: v.to(dtype) for k, v in batch["targets"].items()}
    return batch
