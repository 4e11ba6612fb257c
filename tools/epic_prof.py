import sys, time
sys.path.insert(0, '.')
from slowflow_b200 import ColorImage, Context, Image, synth
from slowflow_b200.api import epic_params_default
w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
im, m, edges = synth.epic_case(w, h, n)
ci = ColorImage.from_array(im)
with Context(0) as ctx:
    fx, fy = Image(w, h), Image(w, h)
    for rep in range(2):
        e = edges.copy()
        t0 = time.perf_counter()
        st = ctx.epic(fx, fy, ci, m, e, None)
        print("call %d: %.1f ms" % (rep, 1e3 * (time.perf_counter() - t0)))
