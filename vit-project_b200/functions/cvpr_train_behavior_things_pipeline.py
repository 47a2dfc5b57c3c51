"""Module name the reference's length-experiment driver imports (LEN:1,
`from functions.cvpr_train_behavior_things_pipeline import run_behavioral_training`) but the reference tree does not
contain.  The driver passes the perturbation / resume keys of the perturbation pipeline
(`perturb_*`, `training_run`, `resume_*`, `previous_training_res_path`, LEN:95-137, 229-253), so the name resolves
to that pipeline here and LEN runs unmodified."""
from functions.new_cvpr_train_behavior_things_pipeline import *  # noqa: F401,F403
from functions.new_cvpr_train_behavior_things_pipeline import run_behavioral_training, train_model  # noqa: F401
