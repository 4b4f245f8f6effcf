#!/bin/bash
# NER fine-tuning with the reference's recipe (scripts/run_ner.sh: lr 5e-6, 5 epochs, batch 32, seq 128).
# DATASET selects the label set: conll2003 | jnlpba | ncbi | bc5cdr
CHECKPOINT=${CHECKPOINT:-results/bert_pretraining/pretrain_ckpts/ckpt_8601.pt}
CONFIG=${CONFIG:-config/bert_large_uncased_config.json}
DATASET=${DATASET:-conll2003}
DATA_DIR=${DATA_DIR:-data/ner/$DATASET}
case "$DATASET" in
    conll2003) LABELS="O B-PER I-PER B-ORG I-ORG B-LOC I-LOC B-MISC I-MISC" ;;
    jnlpba)    LABELS="O B-DNA I-DNA B-RNA I-RNA B-cell_line I-cell_line B-cell_type I-cell_type B-protein I-protein" ;;
    ncbi)      LABELS="O B-Disease I-Disease" ;;
    bc5cdr)    LABELS="O B-Chemical I-Chemical B-Disease I-Disease" ;;
    *) echo "unknown DATASET $DATASET"; exit 1 ;;
esac
python run_ner.py \
    --train_file "$DATA_DIR/train.txt" --val_file "$DATA_DIR/dev.txt" --test_file "$DATA_DIR/test.txt" \
    --labels $LABELS --model_config_file "$CONFIG" --model_checkpoint "$CHECKPOINT" \
    --lr ${LR:-5e-6} --epochs ${EPOCHS:-5} --batch_size ${BATCH:-32} --max_seq_len ${SEQ:-128} --bf16 "$@"
