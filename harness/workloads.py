"""Named workloads: the synthetic clips + x265 options behind the golden traces and the bench.
name -> (bit depth, width, height, frames, seed, reference pool threads, x265 options, dump arrays?)"""

WORKLOADS = {
    # tiny: full arrays dumped, whole-frame estimates only (height < 720 disables coop slices)
    "tiny8": (8, 320, 192, 14, 7, 16, [("preset", "medium"), ("bframes", "3"), ("rc-lookahead", "10")], True),
    "tiny10": (10, 320, 192, 14, 7, 16, [("preset", "medium"), ("bframes", "3"), ("rc-lookahead", "10")], True),
    # odd geometry: lowres width/height not multiples of 8 (rounding, edge CUs), no AQ, no weightp, b-adapt 1
    "odd8": (8, 360, 208, 12, 11, 16, [("preset", "medium"), ("bframes", "2"), ("rc-lookahead", "8"), ("aq-mode", "0"),
                                       ("no-weightp", None), ("b-adapt", "1")], True),
    # BASELINE.json configs[0] reduced to 720p/30 frames for the CPU suite
    "c0_720p": (8, 1280, 720, 30, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "20")], False),
    "c0_720p10": (10, 1280, 720, 24, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "20")], False),
    # small pool: batching auto-disables after the first batch (slicetype.cpp:1256,1296) -> sliced searches
    "pool3_720p": (8, 1280, 720, 24, 99, 3, [("preset", "medium"), ("bframes", "3"), ("rc-lookahead", "12")], False),
    # configs[0] full size: 1080p, --preset medium --bframes 4 --rc-lookahead 20
    "c0_1080p": (8, 1920, 1080, 60, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "20")], False),
    # configs[1]: 1080p, --b-adapt 2 --rc-lookahead 40 with cuTree  (bench workload)
    "c1_1080p": (8, 1920, 1080, 60, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "40"),
                                                ("b-adapt", "2")], False),
    # configs[2]: 4K, --rc-lookahead 40 --bframes 8
    "c2_4k": (8, 3840, 2160, 48, 4321, 16, [("preset", "medium"), ("bframes", "8"), ("rc-lookahead", "40"),
                                            ("b-adapt", "2")], False),
    # configs[3]: 10-bit 4K, --preset slow
    "c3_4k10": (10, 3840, 2160, 32, 4321, 16, [("preset", "slow")], False),
}

DESCRIPTIONS = {
    "c1_1080p": "x265 1.9 lookahead, 1080p 8-bit 4:2:0 synthetic, 60 frames, --preset medium --bframes 4 --b-adapt 2 "
                "--rc-lookahead 40, cuTree/AQ/weightp on (BASELINE.json configs[1])",
    "c2_4k": "x265 1.9 lookahead, 4K 8-bit synthetic, 48 frames, --bframes 8 --b-adapt 2 --rc-lookahead 40 (configs[2])",
    "c0_1080p": "x265 1.9 lookahead, 1080p 8-bit synthetic, 60 frames, --preset medium --bframes 4 --rc-lookahead 20 (configs[0])",
    "c3_4k10": "x265 1.9 lookahead, 4K 10-bit synthetic, 32 frames, --preset slow (configs[3])",
}
