"""Mirror of the parts of ``torchreid.utils`` that sit on the test-time path."""
from . import re_ranking as re_ranking_module
from .re_ranking import re_ranking, re_ranking_dev  # noqa: F401  (the function shadows the submodule name, as in the reference)

__all__ = ['re_ranking', 're_ranking_dev']
