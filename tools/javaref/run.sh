#!/bin/bash
# Pin the oracle against the REAL reference: build libgdx-jbullet from its own sources with a JDK + the libgdx jar, run the
# golden driver over every exported scene and write tests/golden/java_<case>.npz (consumed by tests/test_golden.py).
#   bash tools/javaref/run.sh /path/to/gdx.jar [/path/to/reference/src]
# STAGED: this image has no JDK and no gdx jar (the reference's dependency com.badlogicgames.gdx:gdx, version unpinned
# upstream); nothing here runs during the tests or the bench.
set -euo pipefail
GDX=${1:?usage: run.sh gdx.jar [reference-src]}
SRC=${2:-/root/reference/src}
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=$(mktemp -d)
command -v javac >/dev/null || { echo "no javac on PATH"; exit 2; }
python "$ROOT/tools/javaref/export_scenes.py"
find "$SRC" -name '*.java' > "$OUT/sources.txt"
javac -nowarn -cp "$GDX" -d "$OUT/classes" @"$OUT/sources.txt" "$ROOT/tools/javaref/DumpGolden.java"
for f in "$ROOT"/tests/golden/scenes/*.npz; do
  name=$(basename "$f" .npz)
  java -cp "$GDX:$OUT/classes" DumpGolden "$f" "$ROOT/tests/golden/java_$name.npz"
done
java -version 2>&1 | head -1 > "$ROOT/tests/golden/java_version.txt"
echo "done: python -m pytest tests/test_golden.py -k java"
