"""Drop-in ``HeterModelBaselineWDiffCommStage2`` -- the stage-2 GenComm detector (heterogeneous-agent adaptation stage).

Mirrors ``opencood/models/heter_model_baseline_w_gencomm_stage2.py:31-327``.  At inference it differs from stage 1
(``heter_model_baseline_w_gencomm_stage1.py``) only in: the sampler config key (``args['diffcomm']``, :36), the default
``MessageExtractorv2(128, 2)`` when ``message_extractor`` is absent (:93-96), the ``missing_message`` drop rate
(``rand > 0.1``, :261) and the optional ``trick`` that masks the generated feature with the occupied cells of the
input (:284-285, :293-294).  The frozen-module bookkeeping (``fix_modules``, :45-103,180-185) is training-only.
"""
import torch

from .heter_model_baseline_w_gencomm_stage1 import HeterModelBaselineWGenComm
from .message_extractor import MessageExtractorv2


class HeterModelBaselineWDiffCommStage2(HeterModelBaselineWGenComm):
    GENCOMM_KEY = "diffcomm"
    MISSING_KEEP = 0.1
    CROP_MESSAGE = True        # :236

    def __init__(self, args):
        if 'diffcomm' not in args and 'gencomm' in args:
            # every shipped stage-2 yaml provides ``gencomm:`` while the class reads ``diffcomm`` (KeyError as shipped,
            # SURVEY.md App. B.1): accept both
            args = dict(args, diffcomm=args['gencomm'])
        self.trick = args.get('trick', False)
        super().__init__(args)

    @staticmethod
    def _make_message_extractor(args):
        if 'message_extractor' in args:
            return MessageExtractorv2(args['message_extractor']['in_ch'], args['message_extractor']['out_ch'])
        return MessageExtractorv2(128, 2)

    def _before_gencomm(self, feature):
        if self.trick:
            return torch.any(feature, dim=1).to(torch.uint8).unsqueeze(1)
        return None

    def _after_gencomm(self, pred, state):
        return pred * state if state is not None else pred
