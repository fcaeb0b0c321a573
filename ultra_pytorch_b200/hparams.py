"""Minimal TF-style hyper-parameter string parser with the behaviour the reference's plugins rely on
(ultra/utils/hparams.py:160-259, 418-438): `name=value,name=[a,b]`, types taken from the defaults,
unknown names are reported and ignored (hparams.py:219-224)."""
import re


class HParams(object):
    def __init__(self, **defaults):
        self._types = {}
        for k, v in defaults.items():
            setattr(self, k, v)
            self._types[k] = type(v[0]) if isinstance(v, list) and v else type(v)

    @staticmethod
    def _cast(proto, text):
        text = text.strip().strip("'\"")
        if isinstance(proto, bool):
            return text.lower() in ("true", "1", "t", "yes")
        if isinstance(proto, int):
            return int(float(text))
        if isinstance(proto, float):
            return float(text)
        return text

    def parse(self, values):
        if not values:
            return self
        # split on commas that are not inside brackets
        for item in re.findall(r"[^,\[\]]+=\s*\[[^\]]*\]|[^,\[\]]+=[^,\[\]]*", values):
            name, _, val = item.partition("=")
            name = name.strip()
            if not name:
                continue
            if not hasattr(self, name) or name.startswith("_"):
                print("%s not supported in hparams" % name)
                continue
            cur = getattr(self, name)
            val = val.strip()
            if val.startswith("["):
                elems = [e for e in val.strip("[]").split(",") if e.strip() != ""]
                proto = cur[0] if isinstance(cur, list) and cur else 0
                setattr(self, name, [self._cast(proto, e) for e in elems])
            elif isinstance(cur, list):
                proto = cur[0] if cur else 0
                setattr(self, name, [self._cast(proto, val)])
            else:
                setattr(self, name, self._cast(cur, val))
        return self

    def values(self):
        return {k: getattr(self, k) for k in self._types}

    def to_json(self):
        import json
        return json.dumps(self.values(), sort_keys=True)
