"""Builds the reference's cffi extension module `bliss._bliss` RE-POINTED at this repo's libbliss.so, and stages the
reference's unmodified Python package next to it (oracle/_ref/pyref/bliss/), so that the reference's own
python/bliss/bl_song.py, distance.py, version.py run against the B200 library.

Mirrors reference python/build_bliss.py:21-38 line by line; the differences are exactly the re-pointing:
  - no `sources=`: the analysers are not compiled in, they come from libbliss.so;
  - `libraries=["bliss"]` (+ library_dirs / rpath of bliss_b200/) instead of avformat / avutil / avcodec / fftw3 / swresample;
  - the header handed to cdef is THIS repo's include/bliss.h, filtered by the reference's own rule (lines starting with '#' dropped).
Build container only (needs /root/reference for the package sources); the outputs travel to the GPU box like oracle/_ref/.
The staged .py files are an installed copy of the reference package (as `pip install --target` would make), git-ignored.
"""
import os
import shutil
import sys

from cffi import FFI

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BLISS_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "oracle", "_ref", "pyref")


def main():
    if not os.path.isdir(os.path.join(REF, "python", "bliss")):
        print(f"{REF}/python/bliss absent: keeping the prebuilt oracle/_ref/pyref (if any)")
        return 0
    os.makedirs(os.path.join(OUT, "bliss"), exist_ok=True)
    ffi = FFI()
    ffi.set_source("bliss._bliss",
                   "#include \"bliss.h\"",
                   libraries=["bliss"],
                   library_dirs=[os.path.join(REPO, "bliss_b200")],
                   include_dirs=[os.path.join(REPO, "include")],
                   extra_link_args=["-Wl,-rpath,$ORIGIN/../../../../bliss_b200"],
                   extra_compile_args=["-std=c99"])
    header = ''.join([i for i in open(os.path.join(REPO, "include", "bliss.h"), 'r').readlines()
                      if not i.strip().startswith("#")])
    ffi.cdef(header)
    ffi.compile(tmpdir=OUT)
    for name in os.listdir(os.path.join(REF, "python", "bliss")):  # the unmodified package: __init__, bl_song, distance, version
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(REF, "python", "bliss", name), os.path.join(OUT, "bliss", name))
    for junk in ("bliss/_bliss.c", "bliss/_bliss.o"):
        p = os.path.join(OUT, junk)
        if os.path.exists(p):
            os.remove(p)
    print("built", [f for f in os.listdir(os.path.join(OUT, "bliss"))])
    return 0


if __name__ == "__main__":
    sys.exit(main())
