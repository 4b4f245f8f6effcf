"""SQuAD v1.1 / v2.0: reading, sliding-window featurisation, answer post-processing and EM/F1 scoring.

Parity targets (reference run_squad.py): ``SquadExample`` / ``InputFeatures`` (:61-128),
``read_squad_examples`` (:131-206), ``convert_examples_to_features`` with ``doc_stride`` windows, answer
span refinement and max-context bookkeeping (:209-420), n-best answer extraction incl. the v2 null
answer (:427-568), ``get_final_text`` de-tokenisation alignment (:570-664), ``_get_best_indices`` /
``_compute_softmax`` (:667-699).  The reference shells out to the official ``evaluate-v1.1.py`` and
parses its stdout (:1197-1204); here the same metric (normalised exact match / token F1) is also built
in (:func:`evaluate_predictions`) so an offline box can score without the downloaded script.

Known reference bug not reproduced: the v2 null-prediction loop indexes the null scores with the *last*
example's id instead of each ``qas_id`` (quirk Q20).
"""
from __future__ import annotations

import collections
import json
import math
import re
import string
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

from .tokenization import BasicTokenizer, whitespace_tokenize


# ---------------------------------------------------------------------------
# tokenizer adapter: HF `tokenizers` objects or the pure-Python BertTokenizer
# ---------------------------------------------------------------------------
class TokenizerAdapter:
    def __init__(self, tok):
        self.tok = tok
        self._hf = hasattr(tok, "encode") and hasattr(tok, "token_to_id") and not hasattr(tok, "wordpiece_tokenizer")

    def tokens(self, text: str) -> List[str]:
        if self._hf:
            return self.tok.encode(text, add_special_tokens=False).tokens
        return self.tok.tokenize(text)

    def token_to_id(self, token: str) -> int:
        i = self.tok.token_to_id(token)
        if i is None:
            i = self.tok.token_to_id("[UNK]")
        return i


def _adapt(tok) -> TokenizerAdapter:
    return tok if isinstance(tok, TokenizerAdapter) else TokenizerAdapter(tok)


# ---------------------------------------------------------------------------
# examples
# ---------------------------------------------------------------------------
@dataclass
class SquadExample:
    qas_id: str
    question_text: str
    doc_tokens: List[str]
    orig_answer_text: Optional[str] = None
    start_position: Optional[int] = None
    end_position: Optional[int] = None
    is_impossible: bool = False

    def __repr__(self) -> str:
        s = f"qas_id: {self.qas_id}, question_text: {self.question_text}, doc_tokens: [{' '.join(self.doc_tokens)}]"
        if self.start_position is not None:
            s += f", start_position: {self.start_position}, end_position: {self.end_position}"
        if self.is_impossible:
            s += ", is_impossible: True"
        return s


@dataclass
class InputFeatures:
    unique_id: int
    example_index: int
    doc_span_index: int
    tokens: List[str]
    token_to_orig_map: Dict[int, int]
    token_is_max_context: Dict[int, bool]
    input_ids: List[int]
    input_mask: List[int]
    segment_ids: List[int]
    start_position: Optional[int] = None
    end_position: Optional[int] = None
    is_impossible: Optional[bool] = None


def _is_ws(c: str) -> bool:
    return c in " \t\r\n" or ord(c) == 0x202F


def read_squad_examples(input_file: str, is_training: bool, version_2_with_negative: bool) -> List[SquadExample]:
    with open(input_file, "r", encoding="utf-8") as f:
        data = json.load(f)["data"]
    examples: List[SquadExample] = []
    for entry in data:
        for para in entry["paragraphs"]:
            text = para["context"]
            doc_tokens: List[str] = []
            char_to_word: List[int] = []
            prev_ws = True
            for c in text:
                if _is_ws(c):
                    prev_ws = True
                else:
                    if prev_ws:
                        doc_tokens.append(c)
                    else:
                        doc_tokens[-1] += c
                    prev_ws = False
                char_to_word.append(len(doc_tokens) - 1)
            for qa in para["qas"]:
                start = end = None
                answer_text = None
                impossible = False
                if is_training:
                    if version_2_with_negative:
                        impossible = bool(qa.get("is_impossible", False))
                    if len(qa["answers"]) != 1 and not impossible:
                        raise ValueError("For training, each question should have exactly 1 answer.")
                    if not impossible:
                        ans = qa["answers"][0]
                        answer_text = ans["text"]
                        off = ans["answer_start"]
                        start = char_to_word[off]
                        end = char_to_word[off + len(answer_text) - 1]
                        # skip examples whose answer cannot be recovered from the whitespace tokens
                        actual = " ".join(doc_tokens[start:end + 1])
                        cleaned = " ".join(whitespace_tokenize(answer_text))
                        if actual.find(cleaned) == -1:
                            continue
                    else:
                        start, end, answer_text = -1, -1, ""
                examples.append(SquadExample(qa["id"], qa["question"], doc_tokens, answer_text, start, end, impossible))
    return examples


def _improve_answer_span(doc_tokens: Sequence[str], start: int, end: int, tokenizer: TokenizerAdapter,
                         orig_answer_text: str) -> Tuple[int, int]:
    """Shrink the token span to the tightest one whose word pieces spell the annotated answer."""
    target = " ".join(tokenizer.tokens(orig_answer_text))
    for s in range(start, end + 1):
        for e in range(end, s - 1, -1):
            if " ".join(doc_tokens[s:e + 1]) == target:
                return s, e
    return start, end


def _check_is_max_context(doc_spans, cur_span_index: int, position: int) -> bool:
    """A token that appears in several windows is 'owned' by the window giving it the most context."""
    best_score, best_idx = None, None
    for i, span in enumerate(doc_spans):
        end = span.start + span.length - 1
        if position < span.start or position > end:
            continue
        score = min(position - span.start, end - position) + 0.01 * span.length
        if best_score is None or score > best_score:
            best_score, best_idx = score, i
    return cur_span_index == best_idx


_DocSpan = collections.namedtuple("DocSpan", ["start", "length"])


def convert_examples_to_features(examples: Sequence[SquadExample], tokenizer, max_seq_length: int, doc_stride: int,
                                 max_query_length: int, is_training: bool,
                                 improve_answer_span: bool = True) -> List[InputFeatures]:
    """Sliding-window featurisation (run_squad.py:209-346).

    ``improve_answer_span``: tighten the training span to the word pieces that spell the annotated answer
    ("paris" instead of "paris ." when the document word is "Paris.").  The reference intends this
    (``_improve_answer_span``, run_squad.py:349-383) but compares against ``tokenizer.encode(answer).tokens``, which
    carries [CLS] / [SEP] and therefore never matches: its spans stay at whole whitespace words.  The default here is
    the intended behaviour; ``False`` reproduces the reference's targets bit for bit
    (tests/test_reference_parity.py)."""
    tok = _adapt(tokenizer)
    unique_id = 1000000000
    features: List[InputFeatures] = []
    for ex_idx, ex in enumerate(examples):
        query = tok.tokens(ex.question_text)[:max_query_length]
        tok_to_orig: List[int] = []
        orig_to_tok: List[int] = []
        all_doc: List[str] = []
        for i, word in enumerate(ex.doc_tokens):
            orig_to_tok.append(len(all_doc))
            for sub in tok.tokens(word):
                tok_to_orig.append(i)
                all_doc.append(sub)
        t_start = t_end = None
        if is_training and ex.is_impossible:
            t_start = t_end = -1
        if is_training and not ex.is_impossible:
            t_start = orig_to_tok[ex.start_position]
            t_end = orig_to_tok[ex.end_position + 1] - 1 if ex.end_position < len(ex.doc_tokens) - 1 else len(all_doc) - 1
            if improve_answer_span:
                t_start, t_end = _improve_answer_span(all_doc, t_start, t_end, tok, ex.orig_answer_text)
        max_doc = max_seq_length - len(query) - 3          # [CLS] q [SEP] doc [SEP]
        spans = []
        off = 0
        while off < len(all_doc):
            length = min(len(all_doc) - off, max_doc)
            spans.append(_DocSpan(off, length))
            if off + length == len(all_doc):
                break
            off += min(length, doc_stride)
        for si, span in enumerate(spans):
            tokens = ["[CLS]"] + query + ["[SEP]"]
            seg = [0] * len(tokens)
            t2o: Dict[int, int] = {}
            is_max: Dict[int, bool] = {}
            for i in range(span.length):
                pos = span.start + i
                t2o[len(tokens)] = tok_to_orig[pos]
                is_max[len(tokens)] = _check_is_max_context(spans, si, pos)
                tokens.append(all_doc[pos])
                seg.append(1)
            tokens.append("[SEP]")
            seg.append(1)
            ids = [tok.token_to_id(t) for t in tokens]
            mask = [1] * len(ids)
            pad = max_seq_length - len(ids)
            ids += [0] * pad
            mask += [0] * pad
            seg += [0] * pad
            sp = ep = None
            if is_training and not ex.is_impossible:
                d0, d1 = span.start, span.start + span.length - 1
                if not (t_start >= d0 and t_end <= d1):
                    sp = ep = 0                       # answer not in this window -> point at [CLS]
                else:
                    shift = len(query) + 2
                    sp, ep = t_start - d0 + shift, t_end - d0 + shift
            if is_training and ex.is_impossible:
                sp = ep = 0
            features.append(InputFeatures(unique_id, ex_idx, si, tokens, t2o, is_max, ids, mask, seg, sp, ep,
                                          ex.is_impossible))
            unique_id += 1
    return features


# ---------------------------------------------------------------------------
# post-processing
# ---------------------------------------------------------------------------
RawResult = collections.namedtuple("RawResult", ["unique_id", "start_logits", "end_logits"])
Prediction = collections.namedtuple("Prediction", ["text", "start_logit", "end_logit"])
_Prelim = collections.namedtuple("Prelim", ["start_index", "end_index", "start_logit", "end_logit"])


def _get_best_indices(logits: Sequence[float], n_best_size: int) -> List[int]:
    return [i for i, _ in sorted(enumerate(logits), key=lambda t: t[1], reverse=True)[:n_best_size]]


def _compute_softmax(scores: Sequence[float]) -> List[float]:
    if not scores:
        return []
    m = max(scores)
    e = [math.exp(s - m) for s in scores]
    z = sum(e)
    return [x / z for x in e]


def get_final_text(pred_text: str, orig_text: str, do_lower_case: bool, verbose_logging: bool = False) -> str:
    """Project the word-piece prediction back onto the original (cased, punctuated) text by aligning the
    non-space characters of both strings."""
    def strip_spaces(text):
        chars, mapping = [], collections.OrderedDict()
        for i, c in enumerate(text):
            if c == " ":
                continue
            mapping[len(chars)] = i
            chars.append(c)
        return "".join(chars), mapping

    tok_text = " ".join(BasicTokenizer(do_lower_case=do_lower_case).tokenize(orig_text))
    start = tok_text.find(pred_text)
    if start == -1:
        return orig_text
    end = start + len(pred_text) - 1
    orig_ns, orig_map = strip_spaces(orig_text)
    tok_ns, tok_map = strip_spaces(tok_text)
    if len(orig_ns) != len(tok_ns):
        return orig_text
    tok_s_to_ns = {v: k for k, v in tok_map.items()}
    o_start = o_end = None
    if start in tok_s_to_ns and tok_s_to_ns[start] in orig_map:
        o_start = orig_map[tok_s_to_ns[start]]
    if end in tok_s_to_ns and tok_s_to_ns[end] in orig_map:
        o_end = orig_map[tok_s_to_ns[end]]
    if o_start is None or o_end is None:
        return orig_text
    return orig_text[o_start:o_end + 1]


def get_answer_text(example: SquadExample, feature: InputFeatures, pred: _Prelim, do_lower_case: bool,
                    verbose_logging: bool = False) -> str:
    tok_tokens = feature.tokens[pred.start_index:pred.end_index + 1]
    o0 = feature.token_to_orig_map[pred.start_index]
    o1 = feature.token_to_orig_map[pred.end_index]
    orig_text = " ".join(example.doc_tokens[o0:o1 + 1])
    tok_text = " ".join(" ".join(tok_tokens).replace(" ##", "").replace("##", "").strip().split())
    return get_final_text(tok_text, orig_text, do_lower_case, verbose_logging)


def get_valid_prelim_predictions(start_indices, end_indices, feature: InputFeatures, result: RawResult,
                                 max_answer_length: int) -> List[_Prelim]:
    out = []
    for s in start_indices:
        for e in end_indices:
            if s >= len(feature.tokens) or e >= len(feature.tokens):
                continue
            if s not in feature.token_to_orig_map or e not in feature.token_to_orig_map:
                continue
            if not feature.token_is_max_context.get(s, False):
                continue
            if e < s or e - s + 1 > max_answer_length:
                continue
            out.append(_Prelim(s, e, result.start_logits[s], result.end_logits[e]))
    return out


def match_results(examples, features, results):
    by_id = {r.unique_id: r for r in results}
    feats = sorted(features, key=lambda f: f.unique_id)
    for f in feats:
        if f.unique_id in by_id:
            yield examples[f.example_index], f, by_id[f.unique_id]


def get_answers(examples: Sequence[SquadExample], features: Sequence[InputFeatures], results: Sequence[RawResult], *,
                n_best_size: int = 20, max_answer_length: int = 30, do_lower_case: bool = True,
                version_2_with_negative: bool = False, null_score_diff_threshold: float = 0.0,
                verbose_logging: bool = False):
    """(answers: qas_id -> text, nbest: qas_id -> list of {text, probability, start_logit, end_logit})."""
    preds: Dict[str, List[Prediction]] = collections.defaultdict(list)
    nulls: Dict[str, Tuple[float, float, float]] = {}
    for ex, feat, res in match_results(examples, features, results):
        s_idx = _get_best_indices(res.start_logits, n_best_size)
        e_idx = _get_best_indices(res.end_logits, n_best_size)
        prelim = sorted(get_valid_prelim_predictions(s_idx, e_idx, feat, res, max_answer_length),
                        key=lambda p: p.start_logit + p.end_logit, reverse=True)
        if version_2_with_negative:
            score = res.start_logits[0] + res.end_logits[0]
            if score < nulls.get(ex.qas_id, (float("inf"), 0, 0))[0]:
                nulls[ex.qas_id] = (score, res.start_logits[0], res.end_logits[0])
        seen = {p.text for p in preds[ex.qas_id]}
        cur: List[Prediction] = []
        for p in prelim:
            if len(cur) >= n_best_size:
                break
            text = get_answer_text(ex, feat, p, do_lower_case, verbose_logging)
            if text in seen:
                continue
            seen.add(text)
            cur.append(Prediction(text, p.start_logit, p.end_logit))
        preds[ex.qas_id] += cur
    for ex in examples:                 # every question gets an entry even if no window produced a span
        preds.setdefault(ex.qas_id, [])
    if version_2_with_negative:
        for qid, (_score, sl, el) in nulls.items():
            preds[qid].append(Prediction("", sl, el))
    answers: Dict[str, str] = {}
    nbest: Dict[str, List[dict]] = {}
    for qid, plist in preds.items():
        plist = sorted(plist, key=lambda p: p.start_logit + p.end_logit, reverse=True)[:n_best_size]
        if not plist:
            plist = [Prediction("empty", 0.0, 0.0)]
        probs = _compute_softmax([p.start_logit + p.end_logit for p in plist])
        nbest[qid] = [{"text": p.text, "probability": pr, "start_logit": p.start_logit, "end_logit": p.end_logit}
                      for p, pr in zip(plist, probs)]
        best_non_null = next((p for p in plist if p.text), None)
        if not version_2_with_negative:
            answers[qid] = nbest[qid][0]["text"]
        else:
            null_score = nulls.get(qid, (0.0, 0, 0))[0]
            if best_non_null is None:
                answers[qid] = ""
            else:
                diff = null_score - (best_non_null.start_logit + best_non_null.end_logit)
                answers[qid] = "" if diff > null_score_diff_threshold else best_non_null.text
    return answers, nbest


# ---------------------------------------------------------------------------
# built-in SQuAD metric (same normalisation as the official evaluate-v1.1.py / v2.0)
# ---------------------------------------------------------------------------
def normalize_answer(s: str) -> str:
    s = s.lower()
    s = "".join(ch for ch in s if ch not in set(string.punctuation))
    s = re.sub(r"\b(a|an|the)\b", " ", s)
    return " ".join(s.split())


def f1_score(prediction: str, truth: str) -> float:
    p, t = normalize_answer(prediction).split(), normalize_answer(truth).split()
    if not p or not t:
        return float(p == t)
    common = collections.Counter(p) & collections.Counter(t)
    same = sum(common.values())
    if same == 0:
        return 0.0
    prec, rec = same / len(p), same / len(t)
    return 2 * prec * rec / (prec + rec)


def evaluate_predictions(dataset_file: str, predictions: Dict[str, str]) -> Dict[str, float]:
    with open(dataset_file, "r", encoding="utf-8") as f:
        data = json.load(f)["data"]
    em = f1 = total = 0.0
    for entry in data:
        for para in entry["paragraphs"]:
            for qa in para["qas"]:
                total += 1
                if qa["id"] not in predictions:
                    continue
                golds = [a["text"] for a in qa.get("answers", [])] or [""]
                pred = predictions[qa["id"]]
                em += max(float(normalize_answer(pred) == normalize_answer(g)) for g in golds)
                f1 += max(f1_score(pred, g) for g in golds)
    total = max(total, 1.0)
    return {"exact_match": 100.0 * em / total, "f1": 100.0 * f1 / total}
