"""Local multi-process launcher, one process per device (reference: app/main.py:19-71).

    python -m app.main --fname configs/pretrain/vitl16.yaml --devices cuda:0 cuda:1 ...
"""
import argparse
import logging
import multiprocessing as mp
import os
import pprint

import yaml

parser = argparse.ArgumentParser()
parser.add_argument('--fname', type=str, help='name of config file to load', default='configs.yaml')
parser.add_argument('--devices', type=str, nargs='+', default=['cuda:0'], help='which devices to use on local machine')


def process_main(rank, fname, world_size, devices):
    # each process sees exactly one GPU, as cuda:0
    os.environ['CUDA_VISIBLE_DEVICES'] = str(devices[rank].split(':')[-1])

    logging.basicConfig()
    logger = logging.getLogger()
    logger.setLevel(logging.INFO if rank == 0 else logging.ERROR)
    logger.info(f'called-params {fname}')

    with open(fname, 'r') as y_file:
        params = yaml.load(y_file, Loader=yaml.FullLoader)
    logger.info('loaded params...')
    if rank == 0:
        pprint.PrettyPrinter(indent=4).pprint(params)
        with open(os.path.join(params['logging']['folder'], 'params-pretrain.yaml'), 'w') as f:
            yaml.dump(params, f)

    from app.scaffold import main as app_main
    from src.utils.distributed import init_distributed
    world_size, rank = init_distributed(rank_and_world_size=(rank, world_size))
    logger.info(f'Running... (rank: {rank}/{world_size})')
    app_main(params['app'], args=params)


if __name__ == '__main__':
    args = parser.parse_args()
    num_gpus = len(args.devices)
    mp.set_start_method('spawn')
    procs = [mp.Process(target=process_main, args=(rank, args.fname, num_gpus, args.devices)) for rank in range(num_gpus)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
