"""Drop-in for the one symbol of the reference's `loadsampling` module that the training script uses
(dss2_run.py:20,131): `progressBar`, a generator that prints a terminal progress bar (loadsampling.py:11-37).
The Monte-Carlo load samplers of that module belong to the offline data generation and are out of scope."""


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
