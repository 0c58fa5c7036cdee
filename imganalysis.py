#!/usr/bin/env python
"""Command line of the B200 build — the flags of the reference's `imganalysis.py:8-70`, dispatched to
`pawlikmorph_lsst_b200.main.calcMorphology` like `imganalysis.py:85-108`.

    python imganalysis.py -f images.csv -fo sample -Aall -cas          (README.md:50 of the reference)

File names in the image list are taken as they are (relative to the working directory, like upstream); `-fo` is the OUTPUT
folder (default ./output, helpers.getLocation), `parameters.csv` is written there.

Differences from upstream, all on purpose:
  * `-A` and `-As` are forwarded (upstream parses them and then passes only `-Aall`: imganalysis.py:92-108), i.e. they work
    as main.py:39, :470, :475 intend;
  * `-lif` takes a number (upstream declares it store_true, which makes `-li` unusable: imganalysis.py:24, main.py:325);
  * `--seed` seeds the cleaning noise of maskstarsSEG (unseeded upstream), `--no-clean` skips that step, `-S` adds the
    smoothness statistic (main.py:494 leaves the call commented out), `--device` picks the GPU;
  * `-par` / `-n` are accepted and ignored (the batch is the parallelism); `-sersic` fits the Sersic profile on the device
    (sersic.py); `-cc` runs the star-catalogue flow, `-segmap` / `-starmask` read `segmap_<name>` / `starmask_<name>` from the output
    folder and `-spm` / `-sci` write `segmap_<name>` / `clean_<name>` there, like upstream; `-d` (matplotlib plots) and `-src LSST`
    (data butler) are not covered and raise NotImplementedError.
"""
import sys
from argparse import ArgumentParser


def build_parser():
    parser = ArgumentParser(description="Analyse morphology of galaxies.")
    parser.add_argument("-A", action="store_true", help="Calculate asymmetry parameter")
    parser.add_argument("-As", action="store_true", help="Calculate shape asymmetry parameter")
    parser.add_argument("-Aall", action="store_true", help="Calculate all asymmetries parameters")
    parser.add_argument("-sersic", "--sersic", action="store_true", help="Calculate sersic profile")
    parser.add_argument("-spm", "--savesegmap", action="store_true", help="Save calculated binary pixelmaps (not covered).")
    parser.add_argument("-sci", "--savecleanimg", action="store_true", help="Save cleaned image (not covered).")
    parser.add_argument("-li", "--largeimage", action="store_true", help="Use large cutout for sky background estimation.")
    parser.add_argument("-lif", "--largeimagefactor", type=float, default=1.5,
                        help="Factor to scale cutout image size (--imgsize) so that a better background value can be estimated")
    parser.add_argument("-f", "--file", type=str, help="File which contains list of images, and RA DECS of object in image.")
    parser.add_argument("-fo", "--folder", type=str, default=None, help="Give location for where to save output files/images")
    parser.add_argument("-src", "--imgsource", type=str, default="SDSS", choices=["SDSS", "LSST"],
                        help="Telescope source of the image. Default is SDSS.")
    parser.add_argument("-s", "--imgsize", type=int, default=128, help="Size of image cutout to analyse")
    parser.add_argument("-cc", "--catalogue", type=str, help="Star catalogue of the objects to mask (objID, ra, dec, psfMag_r, type)")
    parser.add_argument("-ns", "--numsig", type=float, default=5., help="Radial extent to which mask out stars if a catalogue is provided.")
    parser.add_argument("-fs", "--filtersize", type=int, default=3, choices=[1, 3, 5, 7, 9, 11, 13, 15], help="Size of kernel for mean filter")
    parser.add_argument("-par", "--parlib", type=str, default="none", choices=["multi", "parsl", "none"],
                        help="Accepted for compatibility; the GPU batch is the parallelism.")
    parser.add_argument("-n", "--cores", type=int, default=1, help="Accepted for compatibility.")
    parser.add_argument("-segmap", "--segmap", action="store_true", help="Precomputed segmentation maps (not covered by the driver).")
    parser.add_argument("-starmask", "--starmask", action="store_true", help="Precomputed star masks (not covered).")
    parser.add_argument("-cas", "--cas", action="store_true", help="Calculate the CAS parameters (Gini, M20, r20, r80, concentration)")
    parser.add_argument("-d", "--diagnostic", action="store_true", help="Diagnostic plots (not covered).")
    parser.add_argument("-S", "--smoothness", action="store_true", help="Also calculate the smoothness S (commented out upstream)")
    parser.add_argument("--seed", type=int, default=0, help="Seed of the cleaning noise (maskstarsSEG)")
    parser.add_argument("--no-clean", action="store_true", help="Skip the cleaning of external sources")
    parser.add_argument("--device", type=str, default="cuda:0")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.segmap and args.catalogue:                                                       # imganalysis.py:75-83
        raise ValueError("Can't provide both a star catalogue and precomputed segmentation map!")
    if args.starmask and args.catalogue:
        raise ValueError("Can't provide both a star catalogue and precomputed star mask file!")
    if args.imgsource != "SDSS":
        raise NotImplementedError("-src LSST: the LSST Butler reader is not part of this build")
    if args.diagnostic:
        raise NotImplementedError("-d: the matplotlib diagnostics are not part of this build")
    import pawlikmorph_lsst_b200 as pm
    imageInfos = pm.helpers.getFiles(args.file)
    outfolder = pm.helpers.getLocation(args.folder)
    pm.main.calcMorphology(imageInfos, outfolder, asymmetry=args.A, shapeAsymmetry=args.As, allAsymmetry=args.Aall,
                           calculateSersic=args.sersic, saveSegMap=args.savesegmap, saveCleanImage=args.savecleanimg,
                           imageSource=args.imgsource, largeImage=args.largeimage, catalogue=args.catalogue,
                           filterSize=args.filtersize, cores=args.cores, parallelLibrary=args.parlib, numberSigmas=args.numsig,
                           segmap=args.segmap, starMask=args.starmask, CAS=args.cas, npix=args.imgsize,
                           largeImgFactor=args.largeimagefactor, smoothness=args.smoothness, clean=not args.no_clean, seed=args.seed,
                           device=args.device)
    return 0


if __name__ == "__main__":
    sys.exit(main())
