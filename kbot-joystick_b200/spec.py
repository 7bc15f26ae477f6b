"""Constants of the path, host side (train.py:22-70, 1206-1269; robot/kbot/metadata.json; robot.mjcf).

The authoritative copy used by the kernels is `kbs_default_params()` in csrc/kbs_api.cu; tests assert that
this module, that function and the oracle's own tables agree.
"""
import math

JOINT_NAMES = (
    "dof_left_hip_pitch_04", "dof_left_hip_roll_03", "dof_left_hip_yaw_03", "dof_left_knee_04", "dof_left_ankle_02",
    "dof_right_hip_pitch_04", "dof_right_hip_roll_03", "dof_right_hip_yaw_03", "dof_right_knee_04",
    "dof_right_ankle_02", "dof_right_shoulder_pitch_03", "dof_right_shoulder_roll_03", "dof_right_shoulder_yaw_02",
    "dof_right_elbow_02", "dof_right_wrist_00", "dof_left_shoulder_pitch_03", "dof_left_shoulder_roll_03",
    "dof_left_shoulder_yaw_02", "dof_left_elbow_02", "dof_left_wrist_00",
)
JOINT_BIAS_DEG = (20.0, 0.0, 0.0, 50.0, -30.0, -20.0, -0.0, 0.0, -50.0, 30.0,
                  0.0, -10.0, 0.0, 90.0, 0.0, 0.0, 10.0, 0.0, -90.0, 0.0)
JOINT_BIASES = tuple(math.radians(d) for d in JOINT_BIAS_DEG)
JOINT_LIMITS = (
    (-1.047198, 2.216568), (-0.20944, 2.268928), (-1.570796, 1.570796), (0.0, 2.70526), (-1.134464, 0.261799),
    (-2.216568, 1.047198), (-2.268928, 0.20944), (-1.570796, 1.570796), (-2.70526, 0.0), (-0.261799, 1.134464),
    (-3.490658, 1.047198), (-1.658063, 0.436332), (-1.671886, 1.671886), (0.0, 2.478368), (-1.37881, 1.37881),
    (-1.047198, 3.490658), (-0.436332, 1.658063), (-1.671886, 1.671886), (-2.478368, 0.0), (-1.37881, 1.37881),
)
KP = (150, 200, 100, 150, 40, 150, 200, 100, 150, 40, 100, 100, 40, 40, 20, 100, 100, 40, 40, 20)
KD = (24.722, 26.387, 3.419, 8.654, 0.990, 24.722, 26.387, 3.419, 8.654, 0.990,
      8.284, 8.257, 0.945, 1.266, 0.295, 8.284, 8.257, 0.945, 1.266, 0.295)
CTRL_LIMIT = (120, 60, 60, 120, 17, 120, 60, 60, 120, 17, 60, 60, 17, 17, 14, 60, 60, 17, 17, 14)

NUM_JOINTS, NUM_COMMANDS, ACTOR_OBS, CRITIC_OBS = 20, 16, 65, 475
BODY_BASE, BODY_LFOOT, BODY_RFOOT, NBODY = 1, 7, 12, 24
SD_GYRO, SD_IMU_QUAT, SD_TOUCH_L, SD_TOUCH_R, NSENSORDATA = 19, 28, 47, 48, 49

REWARD_NAMES = ("linvel", "angvel", "roll_pitch", "base_height", "arm_pos", "single_contact", "no_contact_p",
                "feet_airtime", "feet_orient", "com_distance", "base_accel", "torque")
REWARD_SCALES = (0.2, 0.1, 0.2, 0.2, 0.2, 0.1, 0.1, 1.5, 0.1, 0.05, 0.1, 0.1)
TERMINATION_NAMES = ("bad_z", "not_upright", "episode_length")

# the 21 observations of get_observations (train.py:1155-1204) + the 4 noisy twins consumed by run_actor
OBSERVATION_NAMES = (
    "joint_position", "biased_joint_position", "joint_velocity", "actuator_force", "center_of_mass_inertia",
    "center_of_mass_velocity", "base_position", "base_orientation", "base_linear_velocity", "base_angular_velocity",
    "base_linear_acceleration", "base_angular_acceleration", "actuator_acceleration", "imu_gyro", "left_foot_touch",
    "right_foot_touch", "feet_position", "base_height", "imu_projected_gravity", "projected_gravity", "com_distance",
)
NOISY_OBSERVATION_NAMES = ("noisy_biased_joint_position", "noisy_joint_velocity", "noisy_imu_gyro",
                           "noisy_imu_projected_gravity")

# per-network FLOPs per env-step at hidden size H: 2*(in*H + depth*8*H*H + H*out)
def net_flops(num_in: int, num_out: int, hidden: int = 256, depth: int = 2) -> int:
    return 2 * (num_in * hidden + depth * 8 * hidden * hidden + hidden * num_out)

# mirror_joints (train.py:1574-1582): out[j] = -in[MIRROR_JOINT_SRC[j]] -- legs swapped, arm halves NOT (as written)
MIRROR_JOINT_SRC = (5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19)
