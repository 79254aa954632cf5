"""ORACLE support — test infrastructure only.  The seed-stable synthetic inputs / weights live in the package
(b200seg/utils/synthetic.py: bench.py's product arm needs the generator and must not import oracle/); the oracle side
re-exports them so that the reference, the oracle and the CUDA modules are all fed by ONE generator."""
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent.parent / "medical-image-segmentation-and-classification_b200"
if str(_PKG) not in sys.path:
    sys.path.insert(0, str(_PKG))

from b200seg.utils.synthetic import (IMAGENET_MEAN, IMAGENET_STD, fill_state_dict_, noise_batch,  # noqa: E402,F401
                                     xray_batch)
