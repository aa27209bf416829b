"""Mirror of the parts of ``torchreid.utils`` that sit on the test-time path."""
from .re_ranking import re_ranking  # noqa: F401

__all__ = ['re_ranking']
