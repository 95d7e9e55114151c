"""Static checks of napa_fft3_b200/lisp/ (the sb-alien host layer a maintainer loads instead of napa-fft3.asd).
No Common Lisp implementation exists in this image, so the layer cannot be loaded here; these tests read the
files with a small S-expression reader and check what can be checked without SBCL: balanced forms, the export
list of the reference (package.lisp:28-49), a definition for every export, no redefinition of symbols of locked
packages (the round-1 `sap+` bug), every alien routine present in include/napafft.h, and the .asd components."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LISP = os.path.join(ROOT, "napa_fft3_b200", "lisp")

# /root/reference/package.lisp:28-49, written out so that the test does not read the reference tree
REFERENCE_EXPORTS = """COMPLEX-SAMPLE COMPLEX-SAMPLE-ARRAY REAL-SAMPLE REAL-SAMPLE-ARRAY DIRECTION SCALING WINDOWING
%ENSURE-TWIDDLES %ENSURE-REVERSE GET-REVERSE %ENSURE-FFT GET-FFT GET-WINDOWED-FFT BIT-REVERSE FFT IFFT RECTANGULAR HANN
BLACKMAN* BLACKMAN TRIANGLE BARTLETT GAUSS* GAUSSIAN GAUSSIAN*BARTLETT^X COSINE-SERIES BLACKMAN-HARRIS WINDOW-VECTOR
WINDOWED-FFT WINDOWED-IFFT *RFFT-TWIDDLES* *RIFFT-TWIDDLES* %2RFFT RFFT RIFFT WINDOWED-RFFT WINDOWED-RIFFT""".split()

# symbols of CL / SB-SYS / SB-ALIEN / SB-EXT that the layer uses and must never define itself (package locks)
LOCKED = {"SAP+", "VECTOR-SAP", "INT-SAP", "WITH-PINNED-OBJECTS", "LOAD-SHARED-OBJECT", "DEFINE-ALIEN-ROUTINE",
          "REPLACE", "FILL", "COPY-SEQ", "MAP-INTO", "LENGTH", "POSIX-GETENV", "SAP-REF-DOUBLE"}


def read_forms(text):
    """Top-level forms as nested lists of tokens (strings); raises on unbalanced input."""
    i, n = 0, len(text)
    stack, top = [[]], None

    def emit(x):
        stack[-1].append(x)

    while i < n:
        c = text[i]
        if c == ";":
            while i < n and text[i] != "\n":
                i += 1
        elif text.startswith("#|", i):
            j = text.index("|#", i)
            i = j + 2
        elif c == '"':
            j = i + 1
            while text[j] != '"':
                j += 2 if text[j] == "\\" else 1
            emit(text[i:j + 1]); i = j + 1
        elif text.startswith("#\\", i):
            m = re.match(r"#\\(\w+|.)", text[i:])
            emit(m.group(0)); i += len(m.group(0))
        elif c == "(":
            stack.append([]); i += 1
        elif c == ")":
            assert len(stack) > 1, "unbalanced ')' at offset %d" % i
            done = stack.pop(); emit(done); i += 1
        elif c in " \t\r\n":
            i += 1
        elif c in "'`,":
            i += 2 if text.startswith(",@", i) else 1
        else:
            m = re.match(r"[^\s()\";]+", text[i:])
            emit(m.group(0)); i += len(m.group(0))
    assert len(stack) == 1, "unbalanced '(' (%d forms left open)" % (len(stack) - 1)
    return stack[0]


def files():
    return sorted(f for f in os.listdir(LISP) if f.endswith(".lisp"))


def forms_of(name):
    return read_forms(open(os.path.join(LISP, name)).read())


def walk(form):
    if isinstance(form, list):
        yield form
        for x in form:
            yield from walk(x)


def test_every_file_reads_and_is_balanced():
    assert files(), "no lisp files"
    for f in files() + ["napa-fft3-b200.asd"]:
        forms = forms_of(f)
        assert forms and all(isinstance(x, list) for x in forms), f


def test_asd_components_exist_in_load_order():
    asd = open(os.path.join(LISP, "napa-fft3-b200.asd")).read()
    comps = re.findall(r'\(:file\s+"([^"]+)"\)', asd)
    assert comps[0] == "package" and comps[1] == "ffi"
    assert sorted(c + ".lisp" for c in comps) == files()


def test_exports_cover_the_reference_package():
    pk = forms_of("package.lisp")
    napa = [f for f in pk if f[0].lower() == "defpackage" and f[1] == '"NAPA-FFT"'][0]
    exported = {tok.strip('"') for sub in napa if isinstance(sub, list) and sub[0].lower() == ":export" for tok in sub[1:]}
    missing = [s for s in REFERENCE_EXPORTS if s not in exported]
    assert not missing, missing
    # the implementation package must see them without a prefix
    impl = [f for f in pk if f[0].lower() == "defpackage" and f[1] == '"NAPA-FFT.B200"'][0]
    uses = {tok.strip('"') for sub in impl if isinstance(sub, list) and sub[0].lower() == ":use" for tok in sub[1:]}
    assert {"CL", "SB-ALIEN", "SB-SYS", "NAPA-FFT"} <= uses


DEFINERS = {"defun", "defmacro", "deftype", "defvar", "defparameter", "define-condition", "define-memoized", "defgeneric"}


def definitions():
    out = {}
    for f in files():
        for top in forms_of(f):
            for form in walk(top):
                if form and isinstance(form[0], str) and form[0].lower() in DEFINERS and len(form) > 1 and isinstance(form[1], str):
                    out.setdefault(form[1].upper(), []).append(f)
    return out


def test_every_export_is_defined_once_and_nothing_locked_is_redefined():
    defs = definitions()
    for sym in REFERENCE_EXPORTS + ["FFT-BATCH", "FILTER-CHUNKS", "NAPAFFT-ERROR"]:
        assert sym in defs, "exported %s has no definition" % sym
        assert len(defs[sym]) == 1, "%s defined in %s" % (sym, defs[sym])
    clash = sorted(LOCKED & set(defs))
    assert not clash, "redefines symbols of locked packages: %s" % clash


def test_alien_routines_match_the_header():
    header = open(os.path.join(ROOT, "include", "napafft.h")).read()
    declared = set(re.findall(r"\b(napafft_\w+)\s*\(", header))
    ffi = open(os.path.join(LISP, "ffi.lisp")).read()
    routines = re.findall(r'\(define-alien-routine\s+"(napafft_\w+)"', ffi)
    assert len(routines) >= 8
    for r in routines:
        assert r in declared, r
    # every call site napafft-foo-bar corresponds to a declared routine
    used = set()
    for f in files():
        used |= set(re.findall(r"\(napafft-([a-z0-9-]+)", open(os.path.join(LISP, f)).read()))
    used -= {"error", "error-status", "error-message"}
    declared_lisp = {r[len("napafft_"):].replace("_", "-") for r in routines}
    assert used <= declared_lisp, used - declared_lisp


def test_no_self_recursive_wrappers():
    """a (defun f (...) (f ...)) whose body is a single call of itself is the round-1 sap+ bug"""
    for f in files():
        for top in forms_of(f):
            if top and isinstance(top[0], str) and top[0].lower() == "defun" and len(top) == 4 and isinstance(top[3], list):
                assert top[3][0].upper() != top[1].upper(), (f, top[1])
