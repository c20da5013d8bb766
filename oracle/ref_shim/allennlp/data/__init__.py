from probnmn_clevr_b200.vocabulary import Vocabulary  # noqa: F401  (duck-typed stand-in)
