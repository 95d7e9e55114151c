import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import napa_fft3_b200 as napa
import parity_suite as S
lg = int(sys.argv[1]); io = bool(int(sys.argv[2])); batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 1 << lg
if batch == 1:
    print("fft", S.check_fft(napa._default, n, io, None, None))
else:
    x = np.stack([S.rand_c(i, n) for i in range(batch)])
    y = napa.fft_batch(x.copy(), in_order=io)
    print("batch ok", S.nre(y[batch - 1], np.fft.fft(x[batch - 1]) if io else 0 * y[batch - 1] + y[batch - 1]))
