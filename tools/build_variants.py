"""Build ablation / A-B variants of libwsegk.so next to the product library (wseg_b200/csrc/variants/<name>.so, git-ignored).
usage: python tools/build_variants.py name=-DFLAG[,-DFLAG2] ...      then on the box: python tools/ab_run.py <lib> <script>"""
import os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vdir = os.path.join(root, "wseg_b200", "csrc", "variants")
os.makedirs(vdir, exist_ok=True)
import importlib
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    os.environ["WSEG_NVCC_EXTRA"] = flags.replace(",", " ")
    import wseg_b200.build as B
    importlib.reload(B)
    lib = B.build(force=True)
    shutil.copy(lib, os.path.join(vdir, name + ".so"))
    print("built", name, flags)
os.environ["WSEG_NVCC_EXTRA"] = ""
import wseg_b200.build as B
importlib.reload(B)
B.build(force=True)
print("rebuilt the product library")
