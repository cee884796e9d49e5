"""Configuration constants of the Sheet03 two-stream path -- same names and values as the reference's
`Sheet03/parameters.py:1-46` (so `from parameters import *` code keeps working), plus the knobs this
implementation adds.  tests/test_parameters.py checks the table against tests/golden/reference_parameters.json,
which oracle/make_golden.py dumped from the reference module itself.
"""

# name -> (value, reference line in Sheet03/parameters.py)
_REFERENCE_TABLE = {
    # sampling / loader
    "VIDEO_FRAME_SAMPLE_RATE": (10, 2), "CONVERT": (False, 3), "VIDEO_INPUT_FRAME_COUNT": (3, 4),
    "VIDEO_INPUT_FLOW_COUNT": (10, 5), "SPATIAL_BATCH_SIZE": (60, 6), "TEMPORAL_BATCH_SIZE": (32, 7),
    "NWORKERS_LOADER": (4, 8), "SHUFFLE_LOADER": (True, 9),
    # transforms
    "CROP_SIZE_TF": (224, 10), "HORIZONTAL_FLIP_TF": (True, 11), "NORM_MEANS_TF": ([0.485, 0.456, 0.406], 12),
    "NORM_STDS_TF": ([0.229, 0.224, 0.225], 13), "COLOR_JITTERS": ([0, 0, 0, 0], 21),
    # network / optimisation
    "NACTION_CLASSES": (101, 14), "NEPOCHS": (25, 15), "INITIAL_LR": (0.1, 16), "MOMENTUM_VAL": (0.9, 17),
    "MILESTONES_LR": ([10, 20], 18), "VIDEO_DESCRIPTOR_DIM": (256, 19), "N_FIXED_LAYERS": (5, 20),
    # file-system constants (the absolute paths are the reference authors' machine; kept verbatim)
    "VIDEO_EXTN": (".avi", 24), "FRAME_EXTN": (".jpg", 25),
    "DATA_DIR": ("/media/data/fmthoker/mini-UCF-101", 26),
    "FLOW_DATA_DIR": ("/media/data/fmthoker/mini-ucf101_flow_img_tvl1_gpu", 27),
    "FRAMES_DIR_TRAIN": ("/media/remote_home/va06/VA/Sheet03/mini-UCF-101-frames-train", 28),
    "FRAMES_DIR_TEST": ("/media/remote_home/va06/VA/Sheet03/mini-UCF-101-frames-test", 29),
    "VIDEOLIST_TRAIN": ("/media/remote_home/va06/VA/Sheet03/demoTrain.txt", 30),
    "VIDEOLIST_TEST": ("/media/remote_home/va06/VA/Sheet03/demoTest.txt", 31),
    "ACTIONLABEL_FILE": ("/media/remote_home/va06/VA/Sheet03/classInd.txt", 32),
    "CHECKPOINT_DIR": ("/media/remote_home/va06/VA/Sheet03/checkpoints/", 33),
    "SPATIAL_CKP_FILE": ("spatial_ckp.pth.tar", 34), "SPATIAL_BEST_FILE": ("spatial_best.pth.tar", 35),
    "MOTION_CKP_FILE": ("temporal_ckp.pth.tar", 36), "MOTION_BEST_FILE": ("temporal_best.pth.tar", 37),
    "X_PREFIX_FLOW": ("flow_x_", 38), "Y_PREFIX_FLOW": ("flow_y_", 39),
    "TEMPORAL_TRAIN_CSV_LOC": ("./temporal_train.csv", 40), "TEMPORAL_TEST_CSV_LOC": ("./temporal_test.csv", 41),
    "SPATIAL_TEST_CSV_LOC": ("./spatial_test.csv", 42), "SPATIAL_TRAIN_CSV_LOC": ("./spatial_train.csv", 43),
    "SPATIAL_PERFORMANCE_LOC": ("./spatial_performance.csv", 44),
    "TEMPORAL_PERFORMANCE_LOC": ("./temporal_performance.csv", 45), "SVM_FILE": ("svm_classifier.pkl", 46),
}
globals().update({_k: _v for _k, (_v, _line) in _REFERENCE_TABLE.items()})

# ---- additions of this implementation (not in the reference) -------------------------------------------
N_TEST_SNIPPETS = 25          # test protocol of notes.txt:113-116 / 225-230: 25 equally spaced snippets ...
N_TEST_CROPS = 10             # ... x (4 corners + centre) x (plain, h-flipped)
STREAM_WEIGHT_SPATIAL = 1.0   # late-fusion weights for the class-score average (notes.txt:229-230)
STREAM_WEIGHT_TEMPORAL = 1.0
FLOW_NORM_MEAN = NORM_MEANS_TF[0]   # 1-channel flow images see only mean[0]/std[0] (2018 torchvision zip semantics)
FLOW_NORM_STD = NORM_STDS_TF[0]
GPU_MAX_BATCH = 250           # snippets per internal chunk of the network workspace (one 250-snippet video)

__all__ = list(_REFERENCE_TABLE) + ["N_TEST_SNIPPETS", "N_TEST_CROPS", "STREAM_WEIGHT_SPATIAL", "STREAM_WEIGHT_TEMPORAL",
                                   "FLOW_NORM_MEAN", "FLOW_NORM_STD", "GPU_MAX_BATCH"]
