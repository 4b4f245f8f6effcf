#!/bin/bash
# Build the pre-training datasets end to end:  download -> wikiextractor -> format -> (vocab) -> encode to HDF5.
#   scripts/create_datasets.sh --output data [--nproc 16] [--no-books] [--download] [--format] [--encode] [--encode-type bert|roberta]
# (The reference's version calls a script/flag that does not exist any more, quirk Q25; this one calls
#  utils/encode_data.py --vocab_file.)  Encoding 100 MB of text takes a few minutes per process.
set -e
OUTPUT_DIR=data; NPROC=8; DOWNLOAD=0; FORMAT=0; ENCODE=0; TYPE=bert; VOCAB=""; BOOKS=1
while [[ $# -gt 0 ]]; do
    case $1 in
        -o|--output) OUTPUT_DIR=$2; shift 2 ;;
        -n|-p|--nproc) NPROC=$2; shift 2 ;;
        --no-books) BOOKS=0; shift ;;
        --download) DOWNLOAD=1; shift ;;
        --format) FORMAT=1; shift ;;
        --encode) ENCODE=1; shift ;;
        --encode-type) TYPE=$2; shift 2 ;;
        --vocab) VOCAB=$2; shift 2 ;;
        -h|--help) sed -n 2,6p "$0"; exit 0 ;;
        *) echo "unknown argument $1"; exit 1 ;;
    esac
done
DL=$OUTPUT_DIR/download; FMT=$OUTPUT_DIR/formatted; ENC=$OUTPUT_DIR/encoded
if [[ $DOWNLOAD -eq 1 ]]; then
    python utils/download.py --dir "$DL" --datasets wikicorpus squad weights
    if [[ $BOOKS -eq 1 ]]; then python utils/download.py --dir "$DL" --datasets bookscorpus; fi
fi
if [[ $FORMAT -eq 1 ]]; then
    python -m wikiextractor.WikiExtractor "$DL/wikicorpus/wikicorpus_en.xml" -b 25M --processes "$NPROC" -o "$DL/wikicorpus/data"
    python utils/format.py --dataset wikicorpus --input_dir "$DL/wikicorpus/data" --output_dir "$FMT/wikicorpus" --processes "$NPROC" --shards 256
    if [[ $BOOKS -eq 1 ]]; then
        python utils/format.py --dataset bookscorpus --input_dir "$DL/bookscorpus/download" --output_dir "$FMT/bookscorpus" --processes "$NPROC" --shards 256
    fi
fi
if [[ $ENCODE -eq 1 ]]; then
    VOCAB=${VOCAB:-$DL/weights/uncased_L-24_H-1024_A-16/vocab.txt}
    if [[ ! -f "$VOCAB" ]]; then
        VOCAB=$OUTPUT_DIR/vocab/wordpiece_30k.txt
        python utils/build_vocab.py -i "$FMT" -o "$VOCAB" -s 30000
    fi
    if [[ "$TYPE" == "bert" ]]; then
        python utils/encode_data.py --input_dir "$FMT" --output_dir "$ENC/bert" --vocab_file "$VOCAB" --max_seq_len 128 --next_seq_prob 0.5 --short_seq_prob 0.1 --processes "$NPROC"
        python utils/encode_data.py --input_dir "$FMT" --output_dir "$ENC/bert" --vocab_file "$VOCAB" --max_seq_len 512 --next_seq_prob 0.5 --short_seq_prob 0.1 --processes "$NPROC"
    else
        python utils/encode_data.py --input_dir "$FMT" --output_dir "$ENC/roberta" --vocab_file "$VOCAB" --max_seq_len 512 --next_seq_prob 0 --short_seq_prob 0.1 --processes "$NPROC"
    fi
fi
