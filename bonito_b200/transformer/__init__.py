from bonito_b200.transformer.model import Model, TransformerEncoderLayer, MultiHeadAttention
from bonito_b200.transformer.basecall import basecall
