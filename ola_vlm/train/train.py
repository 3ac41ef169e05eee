"""Drop-in path of ola_vlm/train/train.py (IFT / VPT stages, scripts/train/finetune.sh): the same train()
on the NTP-only classes (LlavaLlamaForCausalLM / LlavaPhi3ForCausalLM, train.py:933-941)."""
from visper_lm_b200.train.entry import ModelArguments, parse_args  # noqa: F401
from visper_lm_b200.train.entry import train as _train


def train(attn_implementation=None, argv=None):
    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    if "--freeze_task_token" not in argv:          # train.py:65 defaults it to True (ola_vlm_train.py:108: False)
        argv += ["--freeze_task_token", "True"]
    return _train(argv, attn_implementation=attn_implementation, distill=False)


if __name__ == "__main__":
    train()
