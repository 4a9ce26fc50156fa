"""Vehicle parameters and host-side lookup tables (computed with numpy exactly as the reference does)."""
import numpy as np

PARAM_KEYS = ['mu', 'C_Sf', 'C_Sr', 'lf', 'lr', 'h', 'm', 'I', 's_min', 's_max', 'sv_min', 'sv_max',
              'v_switch', 'a_max', 'v_min', 'v_max', 'width', 'length']


def default_params():
    """F110Env's default params dict (f110_env.py:132-156), including the v_min = 1e-8 quirk."""
    return {'mu': 1.0489, 'C_Sf': 4.718, 'C_Sr': 5.4562, 'lf': 0.15875, 'lr': 0.17145, 'h': 0.074, 'm': 3.74,
            'I': 0.04712, 's_min': -0.4189, 's_max': 0.4189, 'sv_min': -3.2, 'sv_max': 3.2, 'v_switch': 7.319,
            'a_max': 9.51, 'v_min': 0.00000001, 'v_max': 20.0, 'width': 0.31, 'length': 0.58, 'lidar_max': 30.0}


def params_vector(params):
    return np.array([float(params[k]) for k in PARAM_KEYS], dtype=np.float64)


def theta_tables(theta_dis=2000):
    """ScanSimulator2D.__init__ (laser_models.py:379-381): linspace(0, 2pi, theta_dis) INCLUDING the endpoint."""
    theta_arr = np.linspace(0.0, 2 * np.pi, num=theta_dis)
    return np.sin(theta_arr), np.cos(theta_arr)


def beam_tables(params, num_beams=1080, fov=4.7):
    """RaceCar.__init__ statics (base_classes.py:122-158): beam angles, their cosines, and the distance from the
    lidar to the car's outline along each beam (used by the iTTC test)."""
    scan_ang_incr = fov / (num_beams - 1)
    cosines = np.zeros((num_beams,))
    scan_angles = np.zeros((num_beams,))
    side_distances = np.zeros((num_beams,))
    dist_sides = params['width'] / 2.
    dist_fr = (params['lf'] + params['lr']) / 2.
    for i in range(num_beams):
        angle = -fov / 2. + i * scan_ang_incr
        scan_angles[i] = angle
        cosines[i] = np.cos(angle)
        if angle > 0:
            if angle < np.pi / 2:
                to_side, to_fr = dist_sides / np.sin(angle), dist_fr / np.cos(angle)
            else:
                to_side, to_fr = dist_sides / np.cos(angle - np.pi / 2.), dist_fr / np.sin(angle - np.pi / 2.)
        else:
            if angle > -np.pi / 2:
                to_side, to_fr = dist_sides / np.sin(-angle), dist_fr / np.cos(-angle)
            else:
                to_side, to_fr = dist_sides / np.cos(-angle - np.pi / 2), dist_fr / np.sin(-angle - np.pi / 2)
        side_distances[i] = min(to_side, to_fr)
    return scan_angles, cosines, side_distances
