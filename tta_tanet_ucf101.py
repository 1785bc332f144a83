"""Entry script -- same name, flow and option overrides as the reference's ``tta_tanet_ucf101.py``: for each of the 12
corruptions run ``eval(args)`` (online ViTTA on TANet-R50) and append the top-1 accuracy to the all-result file.
Paths marked "To Specify" are the reference's placeholders; without them (no network / datasets here) the run uses the
synthetic loader of ``vitta_b200.corpus.basics`` and seeded random weights."""
import os

from vitta_b200.corpus.main_eval import eval
from vitta_b200.utils.opts import get_opts
from vitta_b200.utils.utils_ import get_writer_to_all_result

corruptions = ['gauss_shuffled', 'pepper_shuffled', 'salt_shuffled', 'shot_shuffled', 'zoom_shuffled', 'impulse_shuffled',
               'defocus_shuffled', 'motion_shuffled', 'jpeg_shuffled', 'contrast_shuffled', 'rain_shuffled',
               'h265_abr_shuffled']

if __name__ == '__main__':
    args = get_opts()
    args.gpus = [0]
    args.arch = 'tanet'
    args.dataset = 'ucf101'
    # ========================= To Specify ==========================
    args.model_path = os.environ.get('VITTA_MODEL_PATH')                       # .../tanet_ucf.pth.tar
    args.video_data_dir = os.environ.get('VITTA_VIDEO_DIR')
    args.spatiotemp_mean_clean_file = os.environ.get('VITTA_SRC_MEAN')         # list_spatiotemp_mean_*.npy
    args.spatiotemp_var_clean_file = os.environ.get('VITTA_SRC_VAR')           # list_spatiotemp_var_*.npy
    val_vid_list = os.environ.get('VITTA_VID_LIST', '{}.txt')
    result_dir = os.environ.get('VITTA_RESULT_DIR', 'results/{}_{}/tta_{}')
    # ========================= To Specify ==========================
    n_corr = int(os.environ.get('VITTA_N_CORRUPTIONS', len(corruptions)))
    for corr_id, args.corruptions in enumerate(corruptions[:n_corr]):
        print(f'####Starting Evaluation for ::: {args.corruptions} corruption####')
        args.val_vid_list = val_vid_list.format(args.corruptions)
        args.result_dir = result_dir.format(args.arch, args.dataset, args.corruptions)
        epoch_result_list, _ = eval(args=args)
        if int(os.environ.get('RANK', '0')) != 0:
            continue          # torchrun launch (videos sharded over the ranks): the merged accuracy is written once
        if corr_id == 0:
            f_write = get_writer_to_all_result(args)
        f_write.write(' '.join([str(round(float(xx), 3)) for xx in epoch_result_list]) + '\n')
        f_write.flush()
        if corr_id == n_corr - 1:
            f_write.close()
