"""Import-path compatibility with the reference tree (``from src.modeling import BertForPreTraining`` ...).
Every module here only re-exports the implementation in :mod:`bert_pytorch_b200`."""
