"""B200-native implementation of probnmn-clevr's hot path (NMN executor + ProgramGenerator)."""
__version__ = "0.1.0"
