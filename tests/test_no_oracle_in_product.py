"""The oracle is test infrastructure: nothing under transform360_b200/ may import, link, load or execute anything under
oracle/, and bench.py may only do so in its cpu_baseline / reference legs."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


USES_ORACLE = re.compile(r"(import|from|include|CDLL|dlopen).*\boracle\b")


def _touches_oracle(line: str) -> bool:
    return bool(USES_ORACLE.search(line)) or "t360o_" in line or "libt360" in line


def test_the_guard_itself_fires():
    """Positive controls: the pattern must catch what it is there to catch (round 1 shipped it with a doubled backslash,
    which matched nothing)."""
    for bad in ("from oracle import c_oracle", "import oracle.ref_harness as rh", '#include "../../oracle/t360_oracle.h"',
                'ctypes.CDLL("oracle/libt360oracle.so")', "x = t360o_remap(1)"):
        assert _touches_oracle(bad), bad
    for fine in ("# the oracle checks this in tests/", "coracle = 1", "from .handler import load"):
        assert not _touches_oracle(fine), fine


def test_product_sources_never_mention_the_oracle():
    offenders = []
    for f in (ROOT / "transform360_b200").rglob("*"):
        if f.is_file() and f.suffix in {".py", ".cpp", ".cu", ".cuh", ".h"}:
            for i, line in enumerate(f.read_text(errors="ignore").splitlines(), 1):
                if _touches_oracle(line):
                    offenders.append(f"{f.relative_to(ROOT)}:{i}: {line.strip()}")
    assert not offenders, "\n".join(offenders)


def test_product_library_has_no_dependency_on_the_checkers():
    lib = ROOT / "transform360_b200" / "lib" / "libTransform360.so"
    if not lib.exists():
        import pytest
        pytest.skip("library not built")
    needed = subprocess.run(["readelf", "-d", str(lib)], capture_output=True, text=True).stdout
    assert "t360oracle" not in needed and "t360ref" not in needed and "opencv" not in needed.lower()
    assert "libcuda.so" not in needed, "the driver library must not be a link-time dependency (CPU-only hosts load the .so)"


def test_bench_uses_the_oracle_only_in_the_cpu_legs():
    src = (ROOT / "bench.py").read_text()
    body = src[src.index("def main():"):]
    main_imports = [l for l in body.splitlines() if re.search(r"^\s+(from|import) oracle", l)]
    assert not main_imports, "bench.py main() must not import the oracle: " + repr(main_imports)
    ref_fn = src[src.index("def reference_cpu_run"):src.index("def run_reference_arm")]
    assert "from oracle import" in ref_fn
