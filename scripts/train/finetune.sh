#!/bin/bash
# IFT / VPT stage — the reference's scripts/train/finetune.sh (vpt.sh differs in data and output paths only): the PT
# output (a full distilled checkpoint) goes into the NTP-only classes; heads are dropped on load, task tokens stay
# frozen (ola_vlm/train/train.py:65).
PT=${PT:-outputs/pretrain_dsg_VisPer-LM-CLIP-ViT-Llama3-8b}
TOWER=${TOWER:-/ckpt/openai/clip-vit-large-patch14-336}
torchrun --nnodes=1 --nproc-per-node ${GPUS:-8} --master-addr 127.0.0.1 --master-port ${PORT:-29500} \
    -m ola_vlm.train.train_mem \
    --model_name_or_path $PT \
    --version llava_llama_3 \
    --data_path datasets/llava_v1_5_mix665k.json \
    --image_folder datasets/ \
    --vision_tower $TOWER \
    --mm_projector_type mlp2x_gelu \
    --mm_vision_select_layer -2 \
    --mm_use_im_start_end False \
    --mm_use_im_patch_token False \
    --image_aspect_ratio pad \
    --group_by_modality_length True \
    --bf16 True \
    --output_dir outputs/VisPer-LM-CLIP-ViT-Llama3-8b \
    --num_train_epochs 1 \
    --per_device_train_batch_size 16 \
    --gradient_accumulation_steps 1 \
    --save_steps 200 \
    --save_total_limit 3 \
    --learning_rate 2e-5 \
    --weight_decay 0. \
    --warmup_ratio 0.03 \
    --lr_scheduler_type cosine \
    --logging_steps 1 \
    --model_max_length 4096 \
    --dataloader_num_workers 4 \
    --lazy_preprocess True
