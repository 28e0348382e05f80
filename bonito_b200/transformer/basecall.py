"""Transformer models basecall exactly like the CRF models (reference: bonito/transformer/basecall.py re-exports)."""
from bonito_b200.crf.basecall import basecall  # noqa: F401
