"""INFORMATIONAL (not the contract's reference arm): what the reference's GPU path costs on this box
for the part of its step that is 97 % of the FLOPs — HF `LlamaForCausalLM` (the class
ola_llama.py:58 subclasses and llava_llama.py:108 defers to) with flash_attn 2.x
(`attn_implementation="flash_attention_2"`, ola_vlm_train_mem.py:5) + cuBLAS, bf16, frozen weights,
`gradient_checkpointing=True` as in scripts/train/pretrain.sh:52 (and without, for comparison), on
random-init Llama-3-8B weights and an [B, 2048, 4096] embedded sequence that requires grad (the
projector's output).  The CLIP tower, projector, Python splice loop, heads and optimizer the
reference also runs are NOT included, so the real reference step is slower than this number.
Uses only libraries of the image (transformers, flash_attn); none of this repo's kernels."""
import argparse
import json
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--attn", default="flash_attention_2")
    ap.add_argument("--no-ckpt", action="store_true")
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    from transformers import LlamaConfig, LlamaForCausalLM

    dev = torch.device(a.device)
    cfg = LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=a.layers,
                      num_attention_heads=32, num_key_value_heads=8, rope_theta=500000.0,
                      max_position_embeddings=4096, rms_norm_eps=1e-5, tie_word_embeddings=False,
                      attn_implementation=a.attn)
    torch.manual_seed(0)
    with torch.device(dev):
        model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    model.requires_grad_(False)
    model.train()
    if not a.no_ckpt:
        model.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    B, T = a.batch, a.seq
    g = torch.Generator(device="cpu").manual_seed(1)
    labels = torch.randint(0, 128255, (B, T), generator=g)
    labels[:, :622] = -100  # 38 system + 8 + 576 image positions carry no target
    labels = labels.to(dev)

    def step():
        x = torch.randn(B, T, 4096, device=dev, dtype=torch.bfloat16, requires_grad=True)
        out = model(inputs_embeds=x, labels=labels, use_cache=False)
        out.loss.backward()
        return out.loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.steps
    print(json.dumps({"what": "HF LlamaForCausalLM fwd + dgrad-only bwd, frozen weights (library path)",
                      "attn": a.attn, "gradient_checkpointing": not a.no_ckpt, "layers": a.layers,
                      "batch": B, "seq": T, "ms_per_step": round(ms, 1), "samples_per_s": round(B / ms * 1e3, 3),
                      "loss": float(loss), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}),
          flush=True)


if __name__ == "__main__":
    main()
