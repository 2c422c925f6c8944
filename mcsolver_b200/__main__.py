"""python -m mcsolver_b200 parameterfile [workdir]  - the reference's headless entry (mcsolver.loadMC) on the GPU(s); under a
one-process-per-GPU launcher (RANK / WORLD_SIZE / LOCAL_RANK) the (H,T) grid is sharded over the ranks."""
import sys

from . import loadMC

if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    loadMC(sys.argv[1], workdir=sys.argv[2] if len(sys.argv) > 2 else ".")
