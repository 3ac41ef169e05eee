"""Drop-in path of ola_vlm/train/ola_vlm_train_mem.py (the script pretrain.sh launches)."""
from ola_vlm.train.ola_vlm_train import train

if __name__ == "__main__":
    train(attn_implementation="flash_attention_2")
