"""Minimal stand-in for torchgan==0.1.0 (requirements.txt:155 of the reference), which is neither vendored in
/root/reference nor installed.  TEST INFRASTRUCTURE ONLY: it exists so that the reference's own
src/dcgan.py, src/wgan_loss.py and src/betaVAE.py can be imported unmodified when generating golden vectors
(oracle/make_golden.py).  Written from the package's documented behaviour as summarised in SURVEY.md Appendix A;
"parity unpinned" at this boundary -- the reference holds no tests that pin torchgan's results.
"""
from . import losses, models, trainer  # noqa: F401
