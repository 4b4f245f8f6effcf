#!/bin/bash
# SQuAD fine-tuning + evaluation with the reference's recipe (scripts/run_squad.sh: lr 3e-5, 2 epochs, seq 384,
# stride 128, batch 4, 16-bit compute, 1 GPU).
#   scripts/run_squad.sh [CHECKPOINT [OUT_DIR [CONFIG]]]      (positional, as in the reference; or the variables below)
CHECKPOINT=${1:-${CHECKPOINT:-results/bert_pretraining/pretrain_ckpts/ckpt_8601.pt}}
OUT_DIR=${2:-${OUT_DIR:-results/squad}}
CONFIG=${3:-${CONFIG:-config/bert_large_uncased_config.json}}
shift $(( $# < 3 ? $# : 3 ))
SQUAD_DIR=${SQUAD_DIR:-data/download/squad/v1.1}
EPOCHS=${EPOCHS:-2.0}
LR=${LR:-3e-5}
BATCH=${BATCH:-4}
NGPU=${NGPU:-1}
mkdir -p "$OUT_DIR"
LAUNCH="python"
[[ "$NGPU" -gt 1 ]] && LAUNCH="python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc_per_node=$NGPU"
$LAUNCH run_squad.py \
    --init_checkpoint "$CHECKPOINT" --config_file "$CONFIG" --bert_model bert-large-uncased \
    --do_train --train_file "$SQUAD_DIR/train-v1.1.json" --train_batch_size "$BATCH" \
    --do_predict --predict_file "$SQUAD_DIR/dev-v1.1.json" --predict_batch_size "$BATCH" \
    --do_eval --eval_script "$SQUAD_DIR/evaluate-v1.1.py" \
    --do_lower_case --learning_rate "$LR" --num_train_epochs "$EPOCHS" --max_seq_length 384 --doc_stride 128 \
    --seed 42 --output_dir "$OUT_DIR" --fp16 "$@" |& tee "$OUT_DIR/logfile.txt"
