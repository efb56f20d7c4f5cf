"""Exception types used at the drop-in boundary.

Same names as the reference's (autometa/common/exceptions.py) so callers that
catch them keep working; only the two the hot path raises are defined.
"""


class TableFormatError(Exception):
    """A table handed to / read by the k-mer or binning functions is malformed
    (reference: autometa/common/exceptions.py ``TableFormatError``; raised at
    kmers.py:118,564,574 and recursive_dbscan.py:101)."""

    def __init__(self, value):
        self.value = value

    def __str__(self):
        return str(self.value)


class BinningError(Exception):
    """Reference: autometa/common/exceptions.py ``BinningError``."""

    def __init__(self, value):
        self.value = value

    def __str__(self):
        return str(self.value)
