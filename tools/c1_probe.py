import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import egobox_b200 as eg
x = np.sort(np.random.default_rng(42).random((200, 1)) * 25.0, axis=0)
y = (x[:, 0] - 3.5) * np.sin((x[:, 0] - 3.5) / np.pi)
eg.Kriging.params().fit(x[:50], y[:50]).predict_var(x[:10])
for rep in range(3):
    t0 = time.perf_counter(); gp = eg.Kriging.params().fit(x, y); t1 = time.perf_counter()
    xs = np.linspace(0, 25, 100000)[:, None]
    v = gp.predict_var(xs); t2 = time.perf_counter()
    v = gp.predict_var(xs); t3 = time.perf_counter()
    print(json.dumps({"fit_ms": (t1-t0)*1e3, "predict_var_100k_ms": (t2-t1)*1e3, "second": (t3-t2)*1e3}), flush=True)
    gp.close()
