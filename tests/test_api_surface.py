"""Public-name parity with the reference's modules: a user switching over finds the same
entry points.  The names are listed here, not scraped from the reference tree, so the test runs anywhere."""
import importlib
import random

import pytest

SURFACE = {
    "bert_pytorch_b200.models": [
        "BertConfig", "BertLayerNorm", "BertNonFusedLayerNorm", "LinearActivation", "BertEmbeddings", "BertEncoder",
        "BertPooler", "BertModel", "BertForPreTraining", "BertForMaskedLM", "BertForNextSentencePrediction",
        "BertForSequenceClassification", "BertForMultipleChoice", "BertForTokenClassification",
        "BertForQuestionAnswering", "BertPreTrainedModel", "BertPretrainingCriterion", "ACT2FN"],
    "bert_pytorch_b200.optim": [
        "BertAdam", "Lamb", "FusedLAMB", "FusedAdam", "warmup_cosine", "warmup_constant", "warmup_linear", "warmup_poly",
        "LRScheduler", "CosineWarmUpScheduler", "ConstantWarmUpScheduler", "LinearWarmUpScheduler",
        "PolyWarmUpScheduler", "SCHEDULES", "GradScaler"],
    "bert_pytorch_b200.data.tokenization": [
        "convert_to_unicode", "whitespace_tokenize", "BasicTokenizer", "WordpieceTokenizer", "BertTokenizer",
        "get_wordpiece_tokenizer", "get_bpe_tokenizer", "load_vocab"],
    "bert_pytorch_b200.utils.file_utils": [
        "url_to_filename", "filename_to_url", "cached_path", "get_from_cache", "split_s3_path", "s3_request", "s3_etag",
        "s3_get", "http_get", "read_set_from_file", "get_file_extension"],
    "bert_pytorch_b200.finetune_squad": [
        "SquadExample", "InputFeatures", "read_squad_examples", "convert_examples_to_features", "get_answers",
        "get_answer_text", "get_final_text", "get_valid_prelim_predictions", "match_results", "main"],
    "bert_pytorch_b200.pretrain": ["BertPretrainingCriterion", "parse_arguments", "main"],
    "bert_pytorch_b200.data.encode": [
        "TrainingSample", "convert_to_unicode", "get_documents_from_file", "create_samples_from_document",
        "create_samples", "write_samples_to_hdf5", "encode_file"],
}


@pytest.mark.parametrize("module", sorted(SURFACE))
def test_public_names(module):
    m = importlib.import_module(module)
    missing = [n for n in SURFACE[module] if not hasattr(m, n)]
    assert not missing, f"{module} lacks {missing}"


def test_small_helpers(tmp_path):
    from bert_pytorch_b200.data import encode as E
    from bert_pytorch_b200.data.tokenization import convert_to_unicode
    from bert_pytorch_b200.optim import warmup_constant, warmup_cosine, warmup_linear, warmup_poly
    from bert_pytorch_b200.utils import file_utils as F

    assert convert_to_unicode(b"caf\xc3\xa9") == "café" and convert_to_unicode("x") == "x"
    with pytest.raises(ValueError):
        convert_to_unicode(3)
    assert F.split_s3_path("s3://bucket/a/b.txt") == ("bucket", "a/b.txt")
    assert F.get_file_extension("x/Y.TXT") == ".txt" and F.get_file_extension("x/Y.TXT", dot=False, lower=False) == "TXT"
    p = tmp_path / "set.txt"
    p.write_text("a\nb \na\n")
    assert F.read_set_from_file(str(p)) == {"a", "b"}
    # warm-up ramps are x / warmup below the knee
    for f in (warmup_constant, warmup_cosine, warmup_linear, warmup_poly):
        assert f(0.001, 0.002) == pytest.approx(0.5)
    assert warmup_constant(0.5, 0.002) == 1.0 and warmup_linear(1.0, 0.002) == pytest.approx(0.0, abs=1e-6)

    docs = [[[5, 6, 7], [8, 9], [10, 11, 12]], [[20, 21], [22, 23, 24]]]
    samples = E.create_samples_from_document(0, docs, 16, 0.5, 0.1, random.Random(0))
    assert samples and all(isinstance(s, E.TrainingSample) and s.next_seq_ids is not None for s in samples)
    for s in samples:
        ids, special = s.layout(101, 102)
        assert len(ids) <= 16 and ids[0] == 101 and ids[-1] == 102 and len(special) == 3


def test_reference_import_paths():
    """``src.*`` (the reference's package name) resolves to the same objects as the real package."""
    import bert_pytorch_b200 as B
    from bert_pytorch_b200.data.dataset import ShardedPretrainingDataset
    from src.dataset import DistributedSampler, ShardedPretrainingDataset as S2
    from src.modeling import BertConfig, BertForPreTraining
    from src.ner_dataset import NERDataset
    from src.optimization import BertAdam, warmup_linear
    from src.schedulers import PolyWarmUpScheduler
    from src.tokenization import get_wordpiece_tokenizer
    from src.utils import format_step, is_main_process

    assert S2 is ShardedPretrainingDataset and BertForPreTraining is B.models.BertForPreTraining
    assert is_main_process() and callable(format_step) and callable(get_wordpiece_tokenizer)
    assert all(map(callable, (DistributedSampler, BertConfig, NERDataset, BertAdam, warmup_linear, PolyWarmUpScheduler)))
