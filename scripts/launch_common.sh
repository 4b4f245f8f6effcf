#!/bin/bash
# Shared launcher logic for run_pretraining.{sh,sbatch,cobalt}: figure out the node list from the scheduler
# (SLURM / Cobalt / none), build one torchrun command and start it on every node (ssh for remote nodes).
# Expects: CONFIG, DATA, OUTPUT_DIR, optional PRELOAD, BACKEND (nccl|fused) and extra CLI args in "$@".
set -u
PRELOAD="${PRELOAD:-}"
BACKEND="${BACKEND:-fused}"

if [[ -n "${SLURM_NODELIST:-}" ]]; then
    NODEFILE=$(mktemp /tmp/nodefile.XXXX)
    scontrol show hostnames "$SLURM_NODELIST" > "$NODEFILE"
elif [[ -n "${COBALT_NODEFILE:-}" ]]; then
    NODEFILE=$COBALT_NODEFILE
else
    NODEFILE=""
fi
if [[ -z "$NODEFILE" ]]; then
    RANKS=$HOSTNAME; NNODES=1; MASTER_RANK=127.0.0.1
else
    MASTER_RANK=$(head -n 1 "$NODEFILE"); RANKS=$(tr '\n' ' ' < "$NODEFILE"); NNODES=$(< "$NODEFILE" wc -l)
fi

LAUNCHER="python -m torch.distributed.run --nnodes=$NNODES --nproc_per_node=${NPROC_PER_NODE:-auto} --max_restarts ${MAX_RESTARTS:-0} "
if [[ "$NNODES" -eq 1 ]]; then
    LAUNCHER+="--standalone --local-addr 127.0.0.1 "
else
    LAUNCHER+="--rdzv_backend=c10d --rdzv_endpoint=$MASTER_RANK "
    # the peer-memory path is a single NVSwitch domain; across nodes gradients go through NCCL unless the two-level
    # variant is requested (B200_HIER_FUSED=1: node-local fused step + rail-wise NCCL all-reduce)
    if [[ "${B200_HIER_FUSED:-0}" == 1 ]]; then PRELOAD+="export B200_HIER_FUSED=1 ; "; else BACKEND=nccl; fi
fi
CMD="run_pretraining.py --input_dir $DATA --output_dir $OUTPUT_DIR --config_file $CONFIG --backend $BACKEND "
FULL_CMD=" $PRELOAD $LAUNCHER $CMD $* "
echo "Training Command: $FULL_CMD"

RANK=0
for NODE in $RANKS; do
    if [[ "$NODE" == "$HOSTNAME" ]]; then
        echo "Launching rank $RANK on local node $NODE"
        eval "$FULL_CMD" &
    else
        echo "Launching rank $RANK on remote node $NODE"
        ssh "$NODE" "cd $PWD; $FULL_CMD" &
    fi
    RANK=$((RANK+1))
done
wait
