#!/bin/bash
# one-off helper (counterpart of the reference's run.sh): build a 30k WordPiece vocabulary from formatted text
python utils/build_vocab.py -i "${1:-data/formatted}" -o "${2:-data/vocab/wordpiece_30k.txt}" -s 30000 --tokenizer wordpiece
