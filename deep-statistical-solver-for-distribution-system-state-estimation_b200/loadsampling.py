"""Drop-in for the reference's `loadsampling` module: `progressBar` (the one symbol the training script uses, dss2_run.py:20,131;
loadsampling.py:11-37) and the two Monte-Carlo samplers the offline generator can actually run (`samplermontecarlo`,
`samplermontecarlo_normal`, loadsampling.py:75-107), evaluated on the GPU by `dss2_sample_loads` with draws taken from numpy's legacy
global stream like the reference - so seeded runs return the reference's numbers bit for bit.  `kumaraswamymontecarlo`, the grid
samplers and `beta` belong to experiments the repository never calls with valid arguments and are not provided."""
import numpy as np


def progressBar(iterable, prefix="", suffix="", decimals=1, length=100, fill="#", printEnd="\r"):
    items = list(iterable)
    total = len(items)

    def show(done):
        frac = done / float(total) if total else 1.0
        filled = int(length * frac)
        print(f"\r{prefix} |{fill * filled}{'-' * (length - filled)}| {100 * frac:.{decimals}f}% {suffix}", end=printEnd)

    show(0)
    for i, item in enumerate(items):
        yield item
        show(i + 1)
    print()


def samplermontecarlo(LB, UB, numbersamples):
    """loadsampling.py:75-93: LB + rand * (UB - LB); arrays [U] -> [U, n], scalars -> [1, n].  numpy in, numpy out like the reference."""
    from dss2 import sampling
    return sampling.mc_sample(LB, UB, numbersamples, "uniform").cpu().numpy()


def samplermontecarlo_normal(MU, SIG, numbersamples):
    """loadsampling.py:94-107: np.random.normal(MU, SIG) = MU + SIG * gauss; the scalar case returns [1, n] like the reference."""
    from dss2 import sampling
    return sampling.mc_sample(MU, SIG, numbersamples, "normal").cpu().numpy()
