"""File -> JSON throughput of the fused directory driver next to the CPU flow of the reference
(imread -> pipeline -> json.dump, oracle port, all host cores).  python tools/bench_files.py [n_files] [cpu_sample]"""
import json, os, shutil, sys, tempfile, time
import cv2, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200 import synth
from multimodal_biometric_fingerprints_palms_b200.drivers import run_directory


def cpu_one(path):
    from oracle import ref_pipeline as rp
    cv2.setNumThreads(1)
    res = rp.enhance_to_minutiae(cv2.imread(path, cv2.IMREAD_GRAYSCALE))
    with open(path + ".cpu.json", "w") as f:
        json.dump(res["minutiae"], f, indent=2)
    return len(res["minutiae"])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
    cpu_n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    root = tempfile.mkdtemp(prefix="fpb_files_")
    src = os.path.join(root, "in", "cluster_0"); os.makedirs(src)
    base = synth.ridge_batch(64, 320, 240, first_seed=900)
    total = 0
    for i in range(n):
        p = os.path.join(src, f"{i // 10:03d}_{i % 10}.jpg")
        cv2.imwrite(p, base[i % 64], [cv2.IMWRITE_JPEG_QUALITY, 95]); total += os.path.getsize(p)
    out = {"files": n, "jpeg_bytes": total}
    run_directory(os.path.join(root, "in"), os.path.join(root, "warm"), batch=370, write_skeletons=False)
    for tag, sk in (("json_only", False), ("json_and_skeleton_jpegs", True)):
      for lanes in (1, 2):
        for rep in range(2):
            shutil.rmtree(os.path.join(root, "out_" + tag), ignore_errors=True)
            t0 = time.perf_counter()
            st = run_directory(os.path.join(root, "in"), os.path.join(root, "out_" + tag), batch=370, write_skeletons=sk, io_workers=16, lanes=lanes)
            dt = time.perf_counter() - t0
            key = f"{tag}_lanes{lanes}"
            if key not in out or dt < out[key]["seconds"]:
                out[key] = {"files_per_s": n / dt, "seconds": dt, "gpu_decoded": st["gpu_decoded"], "phases": st["seconds"]}
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    files = sorted(os.path.join(src, f) for f in os.listdir(src) if f.endswith(".jpg"))[:cpu_n]
    with ProcessPoolExecutor(cores) as ex:
        list(ex.map(cpu_one, files[:cores]))
        t0 = time.perf_counter(); list(ex.map(cpu_one, files)); dt = time.perf_counter() - t0
    out["cpu_flow"] = {"files_per_s": len(files) / dt, "cores": cores, "sample": len(files)}
    a = json.load(open(files[5] + ".cpu.json"))
    b = json.load(open(os.path.join(root, "out_json_only", "minutiae", "cluster_0", os.path.basename(files[5])[:-4] + "_minutiae.json")))
    out["same_minutiae_as_cpu_flow"] = [(m["x"], m["y"], m["type"]) for m in a] == [(m["x"], m["y"], m["type"]) for m in b]
    print(json.dumps(out))
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
