# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002.
#  *     Distributed under the Boost Software License, Version 1.0. (See
#  *     accompanying file LICENSE_1_0.txt or copy at
#  *     http://www.boost.org/LICENSE_1_0.txt)
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_REPETITION_HPP
# define BOOST_PREPROCESSOR_REPETITION_HPP
#
# include <libint2/boost/preprocessor/repetition/deduce_r.hpp>
# include <libint2/boost/preprocessor/repetition/deduce_z.hpp>
# include <libint2/boost/preprocessor/repetition/enum.hpp>
# include <libint2/boost/preprocessor/repetition/enum_binary_params.hpp>
# include <libint2/boost/preprocessor/repetition/enum_params.hpp>
# include <libint2/boost/preprocessor/repetition/enum_params_with_a_default.hpp>
# include <libint2/boost/preprocessor/repetition/enum_params_with_defaults.hpp>
# include <libint2/boost/preprocessor/repetition/enum_shifted.hpp>
# include <libint2/boost/preprocessor/repetition/enum_shifted_binary_params.hpp>
# include <libint2/boost/preprocessor/repetition/enum_shifted_params.hpp>
# include <libint2/boost/preprocessor/repetition/enum_trailing.hpp>
# include <libint2/boost/preprocessor/repetition/enum_trailing_binary_params.hpp>
# include <libint2/boost/preprocessor/repetition/enum_trailing_params.hpp>
# include <libint2/boost/preprocessor/repetition/for.hpp>
# include <libint2/boost/preprocessor/repetition/repeat.hpp>
# include <libint2/boost/preprocessor/repetition/repeat_from_to.hpp>
#
# endif
