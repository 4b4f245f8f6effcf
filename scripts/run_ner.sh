#!/bin/bash
# NER fine-tuning with the reference's recipe (scripts/run_ner.sh: lr 5e-6, 5 epochs, batch 32, seq 128, cased input).
# DATASET selects the label set (case-insensitive): CoNLL-2003 | JNLPBA | NCBI | BC5CDR
#   DATASET=JNLPBA DATA_DIR=data/ner/JNLPBA CHECKPOINT=results/.../ckpt_8601.pt scripts/run_ner.sh [extra run_ner.py flags]
CHECKPOINT=${CHECKPOINT:-results/bert_pretraining/pretrain_ckpts/ckpt_8601.pt}
CONFIG=${CONFIG:-config/bert_large_uncased_config.json}
DATASET=${DATASET:-CoNLL-2003}
DATA_DIR=${DATA_DIR:-data/download/ner/$DATASET}
UPPERCASE=${UPPERCASE:-true}
case "${DATASET,,}" in
    conll-2003|conll2003) LABELS="O B-PER I-PER B-ORG I-ORG B-MISC I-MISC B-LOC I-LOC" ;;
    jnlpba)               LABELS="O I-DNA B-DNA I-RNA B-RNA I-cell_line B-cell_line I-protein B-protein I-cell_type B-cell_type" ;;
    ncbi|ncbi-disease)    LABELS="O B-Disease I-Disease" ;;
    bc5cdr)               LABELS="O B-Entity I-Entity" ;;
    *) echo "Unknown dataset $DATASET"; exit 1 ;;
esac
KWARGS=""
[[ "$UPPERCASE" == true ]] && KWARGS="--uppercase"
python run_ner.py \
    --train_file "$DATA_DIR/train.txt" --val_file "$DATA_DIR/dev.txt" --test_file "$DATA_DIR/test.txt" \
    --labels $LABELS --model_config_file "$CONFIG" --model_checkpoint "$CHECKPOINT" \
    --lr ${LR:-5e-6} --epochs ${EPOCHS:-5} --batch_size ${BATCH:-32} --max_seq_len ${SEQ:-128} --bf16 $KWARGS "$@"
