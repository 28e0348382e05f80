from bonito_b200.crf.model import Model, SeqdistModel, CTC_CRF, get_stride
from bonito_b200.crf.basecall import basecall
