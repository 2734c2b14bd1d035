"""python-2 `cPickle` stand-in.  The reference opens the SMPL pickle in text mode ('r', batch_smpl.py:34), which
python 3 cannot unpickle from: load() re-opens the same file in binary mode with latin1 strings."""
import pickle as _p

dump, dumps, loads = _p.dump, _p.dumps, _p.loads


def load(f):
    name = getattr(f, "name", None)
    if name is not None and "b" not in getattr(f, "mode", "b"):
        with open(name, "rb") as g:
            return _p.load(g, encoding="latin1")
    return _p.load(f, encoding="latin1")
