"""Drop-in path of ola_vlm/train/train_mem.py (the script finetune.sh launches)."""
from ola_vlm.train.train import train

if __name__ == "__main__":
    train(attn_implementation="flash_attention_2")
