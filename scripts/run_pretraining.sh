#!/bin/bash
# Single-node launcher without a scheduler: PHASE=1|2 CONFIG=... DATA=... OUTPUT_DIR=... scripts/run_pretraining.sh [flags]
PHASE=${PHASE:-1}
CONFIG=${CONFIG:-config/bert_pretraining_phase${PHASE}_config.json}
DATA=${DATA:?set DATA to the directory with the encoded *.hdf5 shards}
OUTPUT_DIR=${OUTPUT_DIR:-results/bert_pretraining}
source "$(dirname "$0")/launch_common.sh" "$@"
