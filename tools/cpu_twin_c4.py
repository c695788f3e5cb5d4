"""CPU side of the 20,000 x 40,000 full-solve parity demonstration (no GPU needed): solve the same
synthetic LP with the binary64 C twin and print the digest of its pivot log, to be compared with
the digest the GPU run printed (tools/parity_full_solves.py c4)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tier_f  # noqa: E402

pos = int(sys.argv[1]) if len(sys.argv) > 1 else 10
threads = int(sys.argv[2]) if len(sys.argv) > 2 else tier_f.lib().tf_max_threads()
m, n = 20000, 40000
A, b, c = tier_f.gen_dense_feasible(m, n, 0, pos, nthreads=threads)
st = tier_f.TierFState(A, b, c, nthreads=threads)
t0 = time.perf_counter()
status, k = st.run()
rec = {"pos_permille": pos, "threads": threads, "status": int(status), "pivots": int(k), "v": float(st.v[0]),
       "wall_s": time.perf_counter() - t0,
       "log_sha256": hashlib.sha256(np.asarray(st.log, dtype=np.int32).tobytes()).hexdigest(),
       "b_sha256": hashlib.sha256(st.b.tobytes()).hexdigest()}
print(json.dumps(rec), flush=True)
json.dump(rec, open(sys.argv[3] if len(sys.argv) > 3 else "/tmp/cpu_twin_c4.json", "w"))
