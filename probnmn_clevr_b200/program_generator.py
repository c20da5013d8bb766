"""``ProgramGenerator`` drop-in (reference: probnmn/models/program_generator.py); see ``seq2seq.py``."""
from .seq2seq import ProgramGenerator, QuestionReconstructor, Seq2SeqBase  # noqa: F401
