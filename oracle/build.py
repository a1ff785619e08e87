"""Compile oracle/oracle.c (plain C, gcc) into oracle/_build/liboracle.so.  Test infrastructure."""
from __future__ import annotations

import hashlib
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "oracle.c"
OUT = HERE / "_build" / "liboracle.so"


def build(force: bool = False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    stamp = OUT.parent / "oracle.sha256"
    digest = hashlib.sha256(SRC.read_bytes()).hexdigest()
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
        return OUT
    gcc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else (shutil.which("gcc") or "gcc")
    subprocess.run([gcc, "-O2", "-fPIC", "-shared", "-std=c11", "-o", str(OUT), str(SRC), "-lm"], check=True)
    stamp.write_text(digest)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
