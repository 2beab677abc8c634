"""String constants of the reference's data dictionaries.

Same names and values as ``InputDataKeys`` / ``OutputDataKeys`` in
``/root/reference/src/data/AddBiomechanicsDataset.py:9-42`` so dicts produced for / by the
reference interoperate unchanged.  The three COM input keys and two output keys that
``TransformerBaseline.forward`` uses but the reference's key classes lack
(``/root/reference/src/models/TransformerBaseline.py:112-114,144-146``) are added here
as extensions; nothing else differs.
"""


class InputDataKeys:
    POS = 'pos'
    VEL = 'vel'
    ACC = 'acc'
    JOINT_CENTERS_IN_ROOT_FRAME = 'jointCentersInRootFrame'
    ROOT_LINEAR_VEL_IN_ROOT_FRAME = 'rootLinearVelInRootFrame'
    ROOT_ANGULAR_VEL_IN_ROOT_FRAME = 'rootAngularVelInRootFrame'
    ROOT_LINEAR_ACC_IN_ROOT_FRAME = 'rootLinearAccInRootFrame'
    ROOT_ANGULAR_ACC_IN_ROOT_FRAME = 'rootAngularAccInRootFrame'
    ROOT_POS_HISTORY_IN_ROOT_FRAME = 'rootPosHistoryInRootFrame'
    ROOT_EULER_HISTORY_IN_ROOT_FRAME = 'rootEulerHistoryInRootFrame'
    # extensions (absent from the reference's class, required by its TransformerBaseline.forward)
    COM_POS = 'comPos'
    COM_VEL = 'comVel'
    COM_ACC = 'comAcc'
    # extensions for the diffusion denoiser (builder-owned): noisy target and timestep
    X_T = 'x_t'
    TIMESTEP = 't'


class OutputDataKeys:
    TAU = 'tau'
    GROUND_CONTACT_WRENCHES_IN_ROOT_FRAME = 'groundContactWrenchesInRootFrame'
    RESIDUAL_WRENCH_IN_ROOT_FRAME = 'residualWrenchInRootFrame'
    CONTACT = 'contact'
    COM_ACC_IN_ROOT_FRAME = 'comAccInRootFrame'
    GROUND_CONTACT_COPS_IN_ROOT_FRAME = 'groundContactCenterOfPressureInRootFrame'
    GROUND_CONTACT_TORQUES_IN_ROOT_FRAME = 'groundContactTorqueInRootFrame'
    GROUND_CONTACT_FORCES_IN_ROOT_FRAME = 'groundContactForceInRootFrame'
    # extensions used by TransformerBaseline.forward (TransformerBaseline.py:142-146)
    COM_ACC = 'comAcc'
    CONTACT_FORCES = 'contactForces'


# model concat order (FeedForwardRegressionBaseline.py:97-108, Groundlink.py:122-133)
MODEL_INPUT_ORDER = (
    InputDataKeys.POS, InputDataKeys.VEL, InputDataKeys.ACC,
    InputDataKeys.ROOT_LINEAR_VEL_IN_ROOT_FRAME, InputDataKeys.ROOT_ANGULAR_VEL_IN_ROOT_FRAME,
    InputDataKeys.ROOT_LINEAR_ACC_IN_ROOT_FRAME, InputDataKeys.ROOT_ANGULAR_ACC_IN_ROOT_FRAME,
    InputDataKeys.JOINT_CENTERS_IN_ROOT_FRAME,
    InputDataKeys.ROOT_POS_HISTORY_IN_ROOT_FRAME, InputDataKeys.ROOT_EULER_HISTORY_IN_ROOT_FRAME,
)

# quantity order of the fused loss kernel and of rows30: [CoP6 | F6 | tau6 | W12]
LOSS_QUANTITIES = (
    OutputDataKeys.GROUND_CONTACT_COPS_IN_ROOT_FRAME,
    OutputDataKeys.GROUND_CONTACT_FORCES_IN_ROOT_FRAME,
    OutputDataKeys.GROUND_CONTACT_TORQUES_IN_ROOT_FRAME,
    OutputDataKeys.GROUND_CONTACT_WRENCHES_IN_ROOT_FRAME,
)
