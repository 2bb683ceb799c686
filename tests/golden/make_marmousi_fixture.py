"""Pack the reference's own Marmousi INPUT fixtures (docs/data/Model_Cp_true.bin, Model_Cp_init_1D.bin,
sourceF_4p5_2_high.bin -- data, not source code) into tests/golden/marmousi_inputs.npz so that the GPU box, which has
no /root/reference, can run the 48-shot geometry of test/TestFWI.jl:6-35 on the real model.

    python tests/golden/make_marmousi_fixture.py        # here, where /root/reference exists

Layout as the reference reads them (test/TestFWI.jl:28-29,47): float32, Julia column-major (nz_pad, nx_pad) =
z fastest; stored here as row-major [z][x] float32 arrays of shape (224, 448).  cs = 0 and rho = 2500 everywhere
(TestFWI.jl:30-31): the acoustic branch (mu_bar = 0) of the elastic kernels on a real model.
"""
import os
import sys

import numpy as np

REF = os.environ.get("FWI_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    d = os.path.join(REF, "docs", "data")
    nz_pad, nx_pad = 224, 448
    rd = lambda f: np.fromfile(os.path.join(d, f), np.float32)
    cp_true = rd("Model_Cp_true.bin").reshape(nx_pad, nz_pad).T.copy()        # column-major (nz, nx) -> [z][x]
    cp_init = rd("Model_Cp_init_1D.bin").reshape(nx_pad, nz_pad).T.copy()
    stf = rd("sourceF_4p5_2_high.bin")
    assert stf.size == 2000 and cp_true.shape == (nz_pad, nx_pad)
    out = os.path.join(HERE, "marmousi_inputs.npz")
    np.savez_compressed(out, cp_true=cp_true, cp_init=cp_init, stf=stf)
    print("written", out, os.path.getsize(out), "bytes; cp range", cp_true.min(), cp_true.max(), "|stf|max", np.abs(stf).max())


if __name__ == "__main__":
    sys.exit(main())
