"""Drop-in `models` package: same import name and registry as the reference
(/root/reference/models/__init__.py:1-21), so `LightningCLI(model_class=models.SRModel,
subclass_mode_model=True)` (main.py:87-93) resolves `--model EDSR|RCAN|RDN|SRCNN` to the
B200-native classes (SRResNet, WDSR: SURVEY §8 f3).  The reference's other models (DDBPN, SRGAN) are outside the
hot path this package accelerates and are not provided."""
from .edsr import EDSR
from .rcan import RCAN
from .rdn import RDN
from .srcnn import SRCNN
from .srmodel import SRModel
from .srresnet import SRResNet
from .wdsr import WDSR

__all__ = ['EDSR', 'RCAN', 'RDN', 'SRCNN', 'SRModel', 'SRResNet', 'WDSR']
