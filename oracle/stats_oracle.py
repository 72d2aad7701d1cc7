"""CPU restatement of the reference's episode-statistics cell `get_average` (Coop-MH-PPO-scalable.py:1550-1675).
TEST INFRASTRUCTURE: only tests/ may import this module.

Parity status: PINNED.  tests/test_ppo_oracle_vs_reference.py::test_stats_oracle_matches_reference_cell executes the
reference's own function text (prints captured, the CO2 helper that needs the unshipped LDV.csv stubbed out) on recorded
evaluation episodes and compares every printed figure.

Inputs in the reference's layout: ep_car [R, C, car_w], ep_ped [R, P, 9], ep_cross [R] (R = rows of back-to-back episodes,
float32).  Returns the quantities the cell prints, with its own names."""
import numpy as np


def get_average(ep_car, ep_ped, ep_cross, nb_car, nb_ped, dt=0.3):
    ep_car, ep_ped, ep_cross = np.asarray(ep_car, np.float32), np.asarray(ep_ped, np.float32), np.asarray(ep_cross, np.float32)
    f32 = np.float32
    temp_cars, all_temp_cars, temp_peds, fin_i, fin_n, decisions, waiting_times, speed_cars, all_decisions, could_stop = [], [], [], [], [], [], [], [], [], []
    R, t = len(ep_cross), 0
    while t + 1 < R:                                                          # PY:1581
        t_init = t_end = t
        for p in range(nb_ped):                                               # PY:1586-1605
            waiting, leave, t = 0.0, 0.0, t_init
            while t + 1 < R and ep_cross[t] == ep_cross[t + 1]:
                if abs(ep_ped[t, p, 3]) == ep_cross[t] and abs(ep_ped[t + 1, p, 3]) == ep_cross[t + 1]:
                    waiting += dt
                if abs(ep_ped[t, p, 3]) == 0 and abs(ep_ped[t + 1, p, 3]) == 0:
                    waiting += dt
                if ep_ped[t, p, 3] * ep_ped[t, p, 8] < ep_cross[t] and ep_ped[t + 1, p, 3] * ep_ped[t + 1, p, 8] >= ep_cross[t + 1]:
                    leave = (t - t_init) * dt
                t += 1
            if leave == 0:
                leave = (t - t_init) * dt
            temp_peds.append(leave); fin_i.append(leave); fin_n.append(leave - waiting); waiting_times.append(waiting)
            t_end = max(t_end, t)
        for i in range(nb_car):                                               # PY:1609-1637
            t, leave, decision = t_init, 0.0, 0.0
            x0, v0, l1 = ep_car[t_init, i, 3], ep_car[t_init, i, 1], ep_car[t_init + 1, i, 4]
            with np.errstate(divide="ignore", invalid="ignore"):
                fin_n.append(float((f32(25.0) - x0) / v0))
            if l1 < 0:
                could_stop.append(1 if (-x0 - (v0 * v0 / f32(8.0) + v0)) > 0 else 0)
            while t + 1 < R and ep_cross[t] == ep_cross[t + 1]:
                decision = float(ep_car[t, i, 4])
                if l1 == 1:
                    speed_cars.append(float(ep_car[t, i, 1]))
                mx = ep_ped[t + 1, :, 2].max()
                if ep_car[t, i, 3] - mx - 25 < 0 and ep_car[t + 1, i, 3] - mx - 25 >= 0:
                    leave = (t - t_init) * dt
                    if l1 == 1:
                        temp_cars.append((t - t_init) * dt)
                t += 1
            if l1 == 1:
                decisions.append(decision)
            if leave == 0:
                leave = (t - t_init) * dt
            all_decisions.append(decision); all_temp_cars.append(leave); fin_i.append(leave)
            t_end = max(t_end, t)
        t = t_end + 1
    std = lambda a: float(np.std(np.asarray(a, np.float64), ddof=1)) if len(a) > 1 else float("nan")     # torch.std: unbiased
    mean = lambda a: float(np.mean(np.asarray(a, np.float64))) if len(a) else float("nan")
    W = nb_car + nb_ped
    FI, FN = np.asarray(fin_i, np.float64).reshape(-1, W), np.asarray(fin_n, np.float64).reshape(-1, W)
    AD = np.asarray(all_decisions).reshape(-1, nb_car)
    n_sc = len(AD)
    out = dict(
        mean_speed=mean(ep_car[:, 0, 1]), sqrt_speed=std(ep_car[:, 0, 1]), mean_acc=mean(np.abs(ep_car[:, 0, 0])), sqrt_acc=std(ep_car[:, 0, 0]),
        mean_speed_p=mean(np.abs(ep_ped[:, 0, 1])), sqrt_speed_p=std(np.abs(ep_ped[:, 0, 1])),
        mean_speed_cars=mean(speed_cars), sqrt_speed_cars=std(speed_cars),
        interaction_cost=mean(FI.max(axis=1) - FN.max(axis=1)), max_finish_compare=mean(FI - FN),
        mean_temp_cars=mean(temp_cars), std_temp_cars=std(temp_cars), mean_all_temp_cars=mean(all_temp_cars), std_all_temp_cars=std(all_temp_cars),
        mean_temp_peds=mean(temp_peds), std_temp_peds=std(temp_peds),
        yield_decision=sum(1 for d in all_decisions if d == 1) / n_sc, go_first_decision=sum(1 for d in all_decisions if d == -1) / n_sc,
        mean_waiting=mean(waiting_times), std_waiting=float(np.std(np.asarray(waiting_times, np.float64))) if waiting_times else float("nan"),
        could_stop=mean(could_stop), n_episodes=n_sc, waiting_times=np.asarray(waiting_times))
    if nb_car == 2:
        s = AD.sum(axis=1)
        out.update(scenario_11=float((s == 2).sum()) / n_sc, scenario_m1m1=float((s == -2).sum()) / n_sc,
                   scenario_m11=float(((AD[:, 0] == -1) & (AD[:, 1] == 1)).sum()) / n_sc, scenario_1m1=float(((AD[:, 0] == 1) & (AD[:, 1] == -1)).sum()) / n_sc)
    return out
