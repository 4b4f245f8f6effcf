"""CoNLL NER dataset (reference src/ner_dataset.py) -> bert_pytorch_b200.data.ner."""
import bert_pytorch_b200.data.ner as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
