"""Minimal stand-in for yacs.config.CfgNode -- only what maskrcnn_benchmark/config/defaults.py needs to build its tree
(attribute-style nested dict, clone, merge_from_list).  yacs is not installed in the build container; this stub is used
by tests/golden/make_golden.py alone, to be able to EXECUTE the reference's ROIBoxHead."""
import copy


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        pass

    def defrost(self):
        pass

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v
