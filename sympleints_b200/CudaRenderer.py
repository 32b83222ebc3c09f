"""CudaRenderer -- the B200 engine behind sympleints' own Renderer plugin surface.

Sits next to ``PythonRenderer`` / ``NumbaRenderer`` / ``FortranRenderer``
(sympleints/Renderer.py:20-165).  ``sympleints.main.run(..., renderers=[CudaRenderer()])`` calls
``render_write(functions, out_dir)`` exactly as for the other renderers (main.py:459-475) and gets
back the path of the rendered module, stored under ``results["cuda"][functions.name]``.

What is rendered per ``Functions`` object (one integral key):

* ``<name>.py`` -- a Python module exposing ``<name> = {(La, Lb[, Lc]): callable}`` with the
  reference's generated-function signature ``f(ax, da, A, bx, db, B[, cx, dc, C], R, result)``
  (templates/py_func.tpl:1-16, py_func_dict.tpl), including the ``La < Lb`` classes that the
  reference serves through its "equivalent" wrappers (Renderer.py:119-137).  Every callable
  evaluates its block ON THE GPU through the C ABI (``si_eval_block``); there is no CPU path.
  The module also exposes ``batched(...)`` helpers that evaluate whole shell sets at once.
* ``<name>.cu`` (3c2e only, next to the .py) -- the class-specialised CUDA device code emitted by
  ``codegen/gen3c2e.py`` for the classes of this Functions object, for inspection; the same
  generator feeds the kernels compiled into ``libsympleints_b200.so``.

The sympy payload of ``Functions`` (``ls_exprs``: CSE replacement lists) is NOT needed: the CUDA
kernels are generated from the recurrence definitions directly.  ``stub_functions()`` builds a
``Functions`` object with empty expressions so that the hours of sympy CSE can be skipped
(hard part 5).

The renderer derives from the reference's ``Renderer`` ABC when ``sympleints`` is importable and
from a structurally identical stand-in otherwise (the GPU box has no reference checkout).
"""
from __future__ import annotations

from pathlib import Path

try:  # the reference package, if it is on sys.path
    from sympleints.Renderer import Renderer as _RefRenderer, RenderedFunction  # type: ignore

    _HAVE_REF = True
except Exception:  # pragma: no cover - exercised on the GPU box
    _HAVE_REF = False

    import abc
    from dataclasses import dataclass
    from typing import Tuple

    @dataclass
    class RenderedFunction:  # sympleints/Renderer.py:13-17
        name: str
        Ls: Tuple[int]
        text: str

    class _RefRenderer(abc.ABC):  # same driver methods as sympleints/Renderer.py:145-165
        def render(self, functions):
            return self.render_module(functions, self.render_functions(functions))

        def write(self, out_dir, name, text):
            fn = (Path(out_dir) / name).with_suffix(self.ext)
            with open(fn, "w") as handle:
                handle.write(text)
            return fn

        def render_write(self, functions, out_dir):
            return self.write(out_dir, functions.name + self.ext, self.render(functions))


# generated-module name -> engine key (sympleints/main.py: names given to Functions(name=...))
NAME_TO_KEY = {
    "ovlp3d": "ovlp",
    "kinetic3d": "kin",
    "dipole3d": "dpm",
    "quadrupole3d": "qpm",
    "diag_quadrupole3d": "dqpm",
    "coulomb3d": "coul",
    "int2c2e3d": "2c2e",
    "int3c2e3d_sph": "3c2e_sph",
    "sph_gto3d": "gto",        # main.py:520-549 (name = ("sph" if sph else "cart") + "_gto3d")
    "cart_gto3d": "gto",
    "prefactors": "prefactor",  # main.py:481-518
    "multipole3d_sph": "multi_sph",  # main.py:752-800
}

_FUNC_TPL = '''
def {fname}({args}, result):
    """{doc}

    Evaluated on the GPU by libsympleints_b200.so (si_eval_block); result is overwritten in place."""
    _engine().eval_block("{key}", {Ls}, {call}, result=result, sph={sph}, normalize="{normalize}")
'''

_MODULE_TPL = '''"""
{header}

CUDA module rendered by sympleints_b200.CudaRenderer for '{name}' (key '{key}').
The callables below have the signature of the reference's generated functions; the arithmetic
runs in hand-written / generated sm_100a CUDA kernels behind the C ABI of libsympleints_b200.so.
"""
import numpy

import sympleints_b200

SPHERICAL = {sph}
NORMALIZE = "{normalize}"
L_MAX = {l_max}
L_AUX_MAX = {l_aux_max}
BOYS_MODULE = {boys!r}

_ENGINE = None


def _boys_table():
    """The Boys table of the module given by --boys-func, if it exposes one (a closure over the
    (nmax+7, 301) table as in sympleints/testing/boys.py:81-91, or a ``table()`` function); the
    reference's table depends on nmax at the 1e-9 level, so parity needs the very same table."""
    if not BOYS_MODULE:
        return None
    import importlib
    try:
        mod = importlib.import_module(BOYS_MODULE)
    except Exception:
        return None
    if hasattr(mod, "table") and callable(mod.table):
        return numpy.asarray(mod.table())
    boys = getattr(mod, "boys", None)

    def dig(fn, depth=0):
        for cell in (getattr(fn, "__closure__", None) or ()):
            v = cell.cell_contents
            if isinstance(v, numpy.ndarray) and v.ndim == 2 and v.shape[1] == 301:
                return v
            if callable(v) and depth < 3:
                t = dig(v, depth + 1)
                if t is not None:
                    return t
        return None

    return dig(boys) if boys is not None else None


def _engine():
    global _ENGINE
    if _ENGINE is None:
        tab = _boys_table()
        lmax = max(L_MAX, 0)
        if tab is not None and tab.shape[0] - 7 < max(2 * L_MAX + L_AUX_MAX, 2 * L_AUX_MAX):
            # same failure mode as the reference (IndexError) is reported per class by the C ABI
            pass
        _ENGINE = sympleints_b200.get_engine(0, lmax, max(L_AUX_MAX, lmax), boys_table=tab)
    return _ENGINE

{funcs}

{name} = {{
{dict_entries}
}}


def batched(shells, aux_shells=None, **kwargs):
    """Whole shell sets at once (the counterpart of sympleints.testing.intor.intor_2c): returns a torch
    tensor on the GPU.  kwargs: R=..., charges=(xyz, q), layout=..., shell_range=..."""
    eng = _engine()
    key = "{key}"
    basis = eng.basis(shells)
    if key == "3c2e_sph":
        return eng.int3c2e(basis, eng.basis(aux_shells), normalize=NORMALIZE,
                           layout=kwargs.get("layout", "full"), shell_range=kwargs.get("shell_range"))
    if key == "coul":
        xyz, q = kwargs["charges"]
        return eng.coul(basis, xyz, q, sph=SPHERICAL, normalize=NORMALIZE)
    if key == "gto":
        return eng.gto(basis, kwargs["points"], sph=SPHERICAL, normalize=NORMALIZE)
    return eng.int1e(key, basis, sph=SPHERICAL, normalize=NORMALIZE, R=kwargs.get("R", (0.0, 0.0, 0.0)))
'''


class CudaRenderer(_RefRenderer):
    ext = ".py"
    language = "CUDA"
    _primitive = False

    def __init__(self, normalize="none", write_cuda_source=True):
        """``normalize`` must repeat the --normalize option given to ``run`` (the Functions payload
        does not carry it; the other renderers get it baked into the sympy expressions)."""
        self.normalize = normalize
        self.write_cuda_source = write_cuda_source

    # ---- the three abstract methods of the reference ABC ---------------------------------------
    def render_function(self, functions, repls, reduced, shape, shape_iter, args, name, L_tots, doc_str=""):
        key = NAME_TO_KEY[functions.name]
        nb = len(L_tots)
        call_args = list(functions.full_args_for_bf_inds(range(nb))) if hasattr(functions, "full_args_for_bf_inds") else list(args)
        has_R = bool(getattr(functions, "with_ref_center", True))
        names = [a for a in call_args]
        if has_R and "R" not in names:
            names.append("R")
        # eval_block(key, Ls, ax, da, A, bx, db, B[, cx, dc, C], R=R)
        prim = names[: 3 * nb]
        call = ", ".join(prim) + (", R=R" if has_R else "")
        if key == "prefactor":
            # the reference's generated prefactors read the module-level names rP / rP2 (they are not arguments)
            call = ", ".join(prim) + ", R=globals()['rP']"
        return _FUNC_TPL.format(fname=name, args=", ".join(names), doc=(doc_str or "").replace('"""', "'''"), key=key,
                                Ls=tuple(L_tots), call=call, sph=bool(functions.spherical) or key == "3c2e_sph",
                                normalize=self.normalize)

    def render_equi_function(self, functions, name, equi_name, equi_inds, shape):
        # The engine accepts any (La, Lb) order, so the "equivalent" class is rendered like a native one.
        Ls = tuple(int(c) for c in equi_name.rsplit("_", 1)[1])
        return self.render_function(functions, None, None, shape, None, functions.full_args, equi_name, Ls,
                                    doc_str=f"See docstring of {name}.")

    def render_functions(self, functions):
        """Same walk as Renderer.render_functions (Renderer.py:74-139) but without touching the sympy
        payload: one callable per native class plus one per swapped (La < Lb) class."""
        rendered = []
        hermi = list(getattr(functions, "hermitian", None) or [])
        for entry in functions.ls_exprs:
            L_tots = tuple(entry[0])
            try:
                doc = functions.doc_func(L_tots)
            except Exception:
                doc = ""
            name = functions.name + "_" + "".join(str(l) for l in L_tots)
            rendered.append(RenderedFunction(name=name, Ls=L_tots,
                                             text=self.render_function(functions, None, None, None, None,
                                                                       functions.full_args, name, L_tots, doc)))
            if len(hermi) == 2 and L_tots[hermi[0]] > L_tots[hermi[1]]:
                sw = list(L_tots)
                sw[hermi[0]], sw[hermi[1]] = sw[hermi[1]], sw[hermi[0]]
                sw = tuple(sw)
                ename = functions.name + "_" + "".join(str(l) for l in sw)
                rendered.append(RenderedFunction(name=ename, Ls=sw,
                                                 text=self.render_equi_function(functions, name, ename, hermi, None)))
        return rendered

    def render_module(self, functions, rendered_funcs, **tpl_kwargs):
        key = NAME_TO_KEY[functions.name]
        entries = "\n".join(f"    {tuple(f.Ls)}: {f.name}," for f in rendered_funcs)
        l_aux = functions.l_aux_max if getattr(functions, "l_aux_max", None) is not None else functions.l_max
        if key in ("2c2e", "3c2e_sph"):
            l_aux = max(max(f.Ls) for f in rendered_funcs)
        l_max = max((max(f.Ls[:2]) for f in rendered_funcs), default=functions.l_max) if key != "2c2e" else 0
        if key == "gto":
            l_aux = l_max
        return _MODULE_TPL.format(header=(functions.header or "").replace('"""', "'''"), name=functions.name, key=key,
                                  sph=bool(functions.spherical) or key == "3c2e_sph", normalize=self.normalize,
                                  l_max=l_max, l_aux_max=l_aux, boys=functions.boys_func,
                                  funcs="\n".join(f.text for f in rendered_funcs), dict_entries=entries)

    def render_write(self, functions, out_dir):
        fn = super().render_write(functions, Path(out_dir))
        if self.write_cuda_source and NAME_TO_KEY.get(functions.name) == "3c2e_sph":
            from .codegen import gen3c2e

            classes = sorted({tuple(sorted(e[0][:2], reverse=True)) + (e[0][2],) for e in functions.ls_exprs})
            (Path(out_dir) / (functions.name + ".cu")).write_text(gen3c2e.generate_file(classes))
        return fn


def stub_functions(name, l_max, l_aux_max=None, spherical=True, ncomponents=1, boys_func=None, three_center=False,
                   one_center=False, with_ref_center=True):
    """A ``Functions``-like payload without sympy expressions (duck-typed; see Functions.py:19-37), so that the
    CudaRenderer can be driven without the reference's expression generation."""
    import itertools as it
    from types import SimpleNamespace

    if three_center:
        classes = [(La, Lb, Lc) for Lb, La in it.combinations_with_replacement(range(l_max + 1), 2)
                   for Lc in range(l_aux_max + 1)]
        args = ["ax", "da", "A", "bx", "db", "B", "cx", "dc", "C", "R"]
    elif one_center:   # gto: f(ax, da, A, R, result)
        classes = [(La,) for La in range(l_max + 1)]
        args = ["ax", "da", "A", "R"]
    else:
        lm = l_max
        classes = [(La, Lb) for Lb, La in it.combinations_with_replacement(range(lm + 1), 2)]
        args = ["ax", "da", "A", "bx", "db", "B"] + (["R"] if with_ref_center else [])
    ns = SimpleNamespace(name=name, l_max=l_max, l_aux_max=l_aux_max, spherical=spherical, ncomponents=ncomponents,
                         boys_func=boys_func, hermitian=[] if one_center else [0, 1],
                         header="stub Functions payload (no sympy expressions)",
                         doc_func=lambda Ls: f"class {Ls}", with_ref_center=with_ref_center, full_args=args,
                         ls_exprs=[(c, ([], [])) for c in classes])
    ns.full_args_for_bf_inds = lambda inds: args
    return ns
