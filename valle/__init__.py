"""Drop-in alias: ``import valle.models`` / ``valle.config`` resolve to the B200-native package, so the reference's
callers (valle/train_model.py:7-9) and tests (tests/test_modules.py:4, tests/test_models_utils.py:4) import it
unchanged.  See INTEGRATION.md."""
import sys

import valle2_b200
import valle2_b200.collate
import valle2_b200.config
import valle2_b200.models
import valle2_b200.models.encodec_pip
import valle2_b200.models.modules
import valle2_b200.models.utils
import valle2_b200.models.valle_ar
import valle2_b200.models.valle_nar

_alias = {
    'valle.collate': valle2_b200.collate,
    'valle.config': valle2_b200.config,
    'valle.models': valle2_b200.models,
    'valle.models.encodec_pip': valle2_b200.models.encodec_pip,
    'valle.models.modules': valle2_b200.models.modules,
    'valle.models.utils': valle2_b200.models.utils,
    'valle.models.valle_ar': valle2_b200.models.valle_ar,
    'valle.models.valle_nar': valle2_b200.models.valle_nar,
}
sys.modules.update(_alias)
collate = valle2_b200.collate
config = valle2_b200.config
models = valle2_b200.models
