"""TEST INFRASTRUCTURE -- a recording stand-in for matplotlib (absent from this image) so that the reference's own
helpers.py can be imported and executed by oracle/gen_golden_eval.py.  Nothing is drawn: every pyplot call is appended to
`pyplot.CALLS` as (name, args, kwargs) so the generator can read back exactly what the reference would have plotted."""
import os  # noqa: F401  (reference helpers.py:4 does `from matplotlib import ..., os`)
from . import pyplot, font_manager  # noqa: F401
