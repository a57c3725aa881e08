"""Model hyper-parameters (reference: permutect/parameters.py:6-40).

Instances are pickled into the saved ``.pt`` under the class path ``permutect.parameters.ModelParameters``
(artifact_model.py:327-342), so that path is part of the file format.  When the reference package is
not importable we register a light alias module under that name, which makes files written by either
implementation loadable by the other.
"""
import importlib.util
import sys
import types
from typing import List


class ModelParameters:
    def __init__(self, read_layers: List[int], self_attention_hidden_dimension: int, num_self_attention_layers: int,
                 info_layers: List[int], aggregation_layers: List[int], num_artifact_clusters: int,
                 calibration_layers: List[int], ref_seq_layers_strings: List[str], dropout_p: float,
                 reweighting_range: float, batch_normalize: bool = False):
        self.read_layers = read_layers
        self.info_layers = info_layers
        self.ref_seq_layer_strings = ref_seq_layers_strings
        self.self_attention_hidden_dimension = self_attention_hidden_dimension
        self.num_self_attention_layers = num_self_attention_layers
        self.aggregation_layers = aggregation_layers
        self.num_artifact_clusters = num_artifact_clusters
        self.calibration_layers = calibration_layers
        self.dropout_p = dropout_p
        self.reweighting_range = reweighting_range
        self.batch_normalize = batch_normalize

    def as_dict(self) -> dict:
        return dict(self.__dict__)


def _reference_importable() -> bool:
    if "permutect.parameters" in sys.modules:
        return not getattr(sys.modules["permutect.parameters"], "_permutect_b200_alias", False)
    try:
        return importlib.util.find_spec("permutect") is not None
    except (ImportError, ValueError):
        return False


def _install_pickle_alias():
    if _reference_importable():
        return
    pkg = sys.modules.get("permutect")
    if pkg is None:
        pkg = types.ModuleType("permutect")
        pkg.__path__ = []
        pkg._permutect_b200_alias = True
        sys.modules["permutect"] = pkg
    mod = types.ModuleType("permutect.parameters")
    mod._permutect_b200_alias = True
    mod.ModelParameters = ModelParameters
    sys.modules["permutect.parameters"] = mod
    pkg.parameters = mod
    ModelParameters.__module__ = "permutect.parameters"


_install_pickle_alias()
