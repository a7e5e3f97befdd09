import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
class Dummy:
    argtypes = None
    restype = None
    def __call__(self, *a): return 0
_CDLL = ctypes.CDLL
class Wrap:
    def __init__(self, p): self._l = _CDLL(p)
    def __getattr__(self, n):
        try: return getattr(self._l, n)
        except AttributeError: return Dummy()
ctypes.CDLL = Wrap
sys.argv = ["gpu_scale_probe.py"] + sys.argv[1:]
P = os.path.join(ROOT, "tools", "gpu_scale_probe.py")
exec(compile(open(P).read(), P, "exec"), {"__file__": P, "__name__": "__main__"})
