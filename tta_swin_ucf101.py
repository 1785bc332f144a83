"""Entry script -- same name, flow and option overrides as the reference's ``tta_swin_ucf101.py`` (Video-Swin-B, T=16,
lr 1e-5, lambda_consis 0.05, momentum 0.05, chosen blocks layers.2 / layers.3 / norm)."""
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
    args.arch = 'videoswintransformer'
    args.dataset = 'ucf101'
    # ========================= To Specify ==========================
    args.model_path = os.environ.get('VITTA_MODEL_PATH')     # swin_base_patch244_window877_pretrain_kinetics400_30epoch_lr3e-5.pth
    args.video_data_dir = os.environ.get('VITTA_VIDEO_DIR')
    args.spatiotemp_mean_clean_file = os.environ.get('VITTA_SRC_MEAN')
    args.spatiotemp_var_clean_file = os.environ.get('VITTA_SRC_VAR')
    val_vid_list = os.environ.get('VITTA_VID_LIST', '{}.txt')
    result_dir = os.environ.get('VITTA_RESULT_DIR', 'results/{}_{}/tta_{}')
    # ========================= To Specify ==========================
    args.clip_length = 16
    args.num_clips = 1
    args.test_crops = 1
    args.frame_uniform = True
    args.frame_interval = 2
    args.scale_size = 224
    args.patch_size = (2, 4, 4)
    args.window_size = (8, 7, 7)
    args.lr = 0.00001
    args.lambda_pred_consis = 0.05
    args.momentum_mvg = 0.05
    args.chosen_blocks = ['module.backbone.layers.2', 'module.backbone.layers.3', 'module.backbone.norm']
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
