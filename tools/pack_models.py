"""Packs the reference's OBJ models into assets/models/*.b2m (binary float32 triangle streams) so the
named scenes can be assembled on machines without /root/reference (the GPU box).  Run once here:
    python tools/pack_models.py [/root/reference/models]
The pack holds exactly the face-vertex stream MeshTriangle consumes (src/Triangle.hpp:99-124)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import b2pt_loader

b2pt = b2pt_loader.load()


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/models"
    os.makedirs(b2pt.ASSET_DIR, exist_ok=True)
    L = b2pt.host_lib()
    for dirpath, _, files in os.walk(src):
        for f in sorted(files):
            if not f.endswith(".obj"):
                continue
            rel = os.path.relpath(os.path.join(dirpath, f), src)
            name = rel[:-4].replace("/", "_") + ".b2m"
            dst = os.path.join(b2pt.ASSET_DIR, name)
            if L.b2pt_host_pack_obj(os.path.join(dirpath, f).encode(), dst.encode()) != 0:
                raise SystemExit(L.b2pt_host_last_error().decode())
            print(rel, "->", name, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
