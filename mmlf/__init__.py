"""Import-compatible alias of the reference package name: ``mmlf.model.feed_forward`` etc. resolve to ``mmlf_b200``, so
``python -m mmlf.train.cli`` / ``python -m mmlf.validate.cli`` and ``from mmlf.model.feed_forward import FeedForward``
work unchanged on the B200 path."""
