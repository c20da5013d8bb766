class CfgNode(dict):
    """attribute-style dict; enough for `probnmn/config.py` to be imported (it is never instantiated here)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, *a, **k):
        raise NotImplementedError

    def merge_from_list(self, *a, **k):
        raise NotImplementedError

    def freeze(self):
        pass
