import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    import numpy as np, torch
    import tie_cases
    from labelany3d_b200 import ops
    pc = tie_cases.cases()[sys.argv[1]]
    for method in ("convex_hull", "sweep", "pca"):
        rec = ops.fit_points(torch.as_tensor(pc).cuda(), torch.tensor([0, len(pc)]).cuda(), None, None, None, method, 36)
        torch.cuda.synchronize()
        print(sys.argv[1], method, "yaw", float(rec[0, 39]), "status", float(rec[0, 41]), flush=True)
else:
    import tie_cases
    for name in sorted(tie_cases.cases()):
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=25)
            print(p.stdout.strip() or p.stderr[-300:], flush=True)
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", name, (e.stdout or b"")[-300:], flush=True)
