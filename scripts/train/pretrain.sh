#!/bin/bash
# PT stage on 8 B200s — the reference's scripts/train/pretrain.sh with its own flags; only the launcher line changes
# (torchrun, one process per GPU over NCCL, instead of the deepspeed launcher; ZeRO-2 is the trainer's own).
# Every path must be local (no network): an HF Llama-3 / Phi-3 directory, an HF CLIP vision directory (or
# CLIP-convnext_xxlarge-res768 with an open_clip checkpoint), the three teacher checkpoints.
LLM=${LLM:-/ckpt/Meta-Llama-3-8B-Instruct}
TOWER=${TOWER:-/ckpt/openai/clip-vit-large-patch14-336}
torchrun --nnodes=1 --nproc-per-node ${GPUS:-8} --master-addr 127.0.0.1 --master-port ${PORT:-29500} \
    -m ola_vlm.train.ola_vlm_train_mem \
    --model_name_or_path $LLM \
    --version llava_llama_3 \
    --mode gen-depth-seg \
    --layer_indices d18-20_s10-18_g12-20 \
    --num_task_tokens 8 \
    --loss_weights d0.5_s0.5_g0.5 \
    --contrastive_loss_weight 0.3 \
    --image_generator ${GEN_TEACHER:-/ckpt/stable-diffusion-2-1-unclip} \
    --image_segmentor ${SEG_TEACHER:-/ckpt/oneformer_coco_swin_large} \
    --depth_estimator ${DEPTH_TEACHER:-/ckpt/depth_anything_v2_vitl.pth} \
    --data_path datasets/LLaVA-Pretrain/blip_laion_cc_sbu_558k.json \
    --image_folder datasets/LLaVA-Pretrain/images \
    --vision_tower $TOWER \
    --mm_projector_type mlp2x_gelu \
    --tune_mm_mlp_adapter True \
    --mm_vision_select_layer -2 \
    --mm_use_im_start_end False \
    --mm_use_im_patch_token False \
    --bf16 True \
    --output_dir outputs/pretrain_dsg_VisPer-LM-CLIP-ViT-Llama3-8b \
    --num_train_epochs 1 \
    --per_device_train_batch_size 32 \
    --gradient_accumulation_steps 1 \
    --save_steps 200 \
    --save_total_limit 3 \
    --learning_rate 1e-3 \
    --weight_decay 0. \
    --warmup_ratio 0.03 \
    --lr_scheduler_type cosine \
    --logging_steps 1 \
    --model_max_length 4096 \
    --dataloader_num_workers 4 \
    --lazy_preprocess True
