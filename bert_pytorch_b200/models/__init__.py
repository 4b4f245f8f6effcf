from ..config import BertConfig  # noqa: F401
from .modeling import (  # noqa: F401
    ACT2FN, BertLayerNorm, BertNonFusedLayerNorm, LinearActivation, BertEmbeddings, BertEncoder, BertLayer, BertPooler,
    BertModel, BertForPreTraining, BertForMaskedLM, BertForNextSentencePrediction,
    BertForSequenceClassification, BertForMultipleChoice, BertForTokenClassification,
    BertForQuestionAnswering, BertPreTrainedModel, BertPretrainingCriterion,
    load_tf_weights_in_bert, bias_gelu, bias_gelu_training, bias_tanh, gelu, swish, count_parameters,
    CONFIG_NAME, WEIGHTS_NAME)
from .arena import ParamArena  # noqa: F401
