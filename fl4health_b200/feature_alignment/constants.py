"""Config keys and enums of the tabular feature-alignment protocol (parity: ``fl4health/feature_alignment/constants.py``)."""

from __future__ import annotations

from enum import Enum

from sklearn.feature_extraction.text import CountVectorizer, HashingVectorizer, TfidfTransformer, TfidfVectorizer

# what a text column may be vectorised with (type alias used by the text column transformers)
TextFeatureTransformer = CountVectorizer | TfidfTransformer | TfidfVectorizer | HashingVectorizer

# server -> client config keys
SOURCE_SPECIFIED = "source_specified"  # has the server fixed the "source of truth" schema yet?
FEATURE_INFO = "feature_info"  # the JSON-encoded schema
# client -> server properties used to size the global model
INPUT_DIMENSION = "input_dimension"
OUTPUT_DIMENSION = "output_dimension"
CURRENT_SERVER_ROUND = "current_server_round"


class FeatureType(Enum):
    NUMERIC = "numeric"
    BINARY = "binary"
    STRING = "string"
    ORDINAL = "ordinal"
    CATEGORICAL_INDICATOR = "categorical_indicator"


FEATURE_TYPES = [FeatureType.NUMERIC, FeatureType.BINARY, FeatureType.STRING, FeatureType.ORDINAL]

FEATURE_INDICATOR_ATTR = "indicator_of"
FEATURE_MAPPING_ATTR = "mapping"
FEATURE_TYPE_ATTR = "type_"
FEATURE_TARGET_ATTR = "target"
FEATURE_META_ATTR_DEFAULTS = {FEATURE_TARGET_ATTR: False, FEATURE_INDICATOR_ATTR: None, FEATURE_MAPPING_ATTR: None}
FEATURE_META_ATTRS = [FEATURE_TYPE_ATTR, FEATURE_TARGET_ATTR, FEATURE_INDICATOR_ATTR, FEATURE_MAPPING_ATTR]
MISSING_CATEGORY = "null_category"
ORDINAL_MAX_CATEGORIES = 20
